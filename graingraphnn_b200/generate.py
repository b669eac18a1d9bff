"""Scalable periodic grain-graph generator: what `graph_trajectory.py --mode=generate` produces,
in vectorised numpy, for domains from the reference's own sizes (lxd 40 / 120 / 240: equal to the reference array for
array) up to the 10^5-10^6-grain configurations the reference's O(N^2) bookkeeping and (lxd/0.08)^2 raster cannot reach.

What is followed, with the reference lines (all `file:line` relative to the reference root):
  * seeds: jittered hexagonal lattice + its 8 periodic images, the same `np.random` stream
    (graph_datastruct.py:118-160, :259-260, :271);
  * `scipy.spatial.Voronoi` of those points (:353) — the third-party routine the reference calls, same input, same output;
  * regions -> vertices (first-seen numbering of the coordinates rounded to 4 decimals), grains (first-seen numbering of
    the vertex sets), vertex -> 3 grains, quadruple splitting (:364-461);
  * `update(init=True)` (:654-757): vertices of a grain in joint2vertex order, periodic chain unwrap, lower-bound shift,
    centre = np.mean, counter-clockwise sort, directed joint-joint edges grain by grain;
  * orientations from the same random stream (:292-305);
  * `form_states_tensor(0)` (graph_trajectory.py:901-1005) + `form_gradient(None, ...)` / `append_history([])`
    (graph_datastruct.py:978-1011, :1035-1036): feature / edge / edge-length / mask arrays of the pickled HeteroGraph;
  * the loader + patch scaling of the rollout driver (data_loader.py:113-162, test.py:29-55, :310-312): `model_inputs`.

What differs by design:
  * the grain `area` feature is a pixel count of the PIL raster in the reference (graph_datastruct.py:553-610, :287-288).
    `area='raster'` hands the polygons to `area_counts_fn` (tests pass the reference's own raster restated on PIL);
    `area='polygon'` (default) uses the exact polygon area in the same units — the scalable choice;
  * beyond lxd = 240 a grain edge is no longer long against the 1e-4 rounding of the vertex coordinates (the reference
    merges distinct vertices there and stops being a trivalent tiling); `decimals=None` keeps 4 decimals up to lxd 240 and adds
    one per factor of 10 in domain size above it.
Host-side data tool (numpy + scipy, like the reference's); nothing here runs per rollout step.
"""
import math

import numpy as np

ET_GJ, ET_JG, ET_JJ = ('grain', 'push', 'joint'), ('joint', 'pull', 'grain'), ('joint', 'connect', 'joint')
EPS = 1e-12
PATCH, MESH = 40, 0.08


# ------------------------------------------------------------------------------------------------ seeds
def lattice_points(dx, noise):
    """hexagonal_lattice(dx, noise, 'periodic') (graph_datastruct.py:118-160) without the Python loops: the in-domain seeds and,
    for each, itself followed by its 8 periodic images in the reference's order.  Consumes the same random numbers."""
    rows, cols = int(1 / dx) + 1, int(1 / dx)
    shiftx, shifty = 0.1 * dx, 0.25 * dx
    rand = np.random.multivariate_normal(mean=np.zeros(2), cov=np.eye(2) * noise, size=rows * cols * 5)
    row = np.repeat(np.arange(rows * 2), cols)
    col = np.tile(np.arange(cols), rows * 2)
    count = np.arange(1, rows * 2 * cols + 1)                       # the reference increments before it indexes
    x = ((col + (0.5 * (row % 2))) * np.sqrt(3)) * dx + shiftx
    y = row * 0.5 * dx + shifty
    x = x + rand[count, 0]
    y = y + rand[count, 1]
    inb = (x >= -EPS) & (x <= 1 + EPS) & (y >= -EPS) & (y <= 1 + EPS)
    x, y = x[inb], y[inb]
    ox = np.array([0, 1, -1, 0, 0, 1, -1, -1, 1], dtype=np.float64)
    oy = np.array([0, 0, 0, 1, -1, 1, -1, 1, -1], dtype=np.float64)
    pts = np.stack([x[:, None] + ox[None, :], y[:, None] + oy[None, :]], axis=-1).reshape(-1, 2)
    return pts, np.stack([x, y], axis=1)


# ------------------------------------------------------------------------------------------------ ragged helpers
def _ragged(lists):
    """list of lists -> (flat int64 array, offsets [n + 1])"""
    lens = np.fromiter((len(r) for r in lists), dtype=np.int64, count=len(lists))
    off = np.zeros(len(lists) + 1, dtype=np.int64)
    np.cumsum(lens, out=off[1:])
    flat = np.fromiter((v for r in lists for v in r), dtype=np.int64, count=int(off[-1]))
    return flat, off


def _first_seen_ids(keys):
    """Number the rows of `keys` [n, k] by first appearance: (id of every row, row index of every id's first appearance)."""
    _, first, inv = np.unique(keys, axis=0, return_index=True, return_inverse=True)
    inv = inv.reshape(-1)
    order = np.argsort(first, kind='stable')                         # unique rows by first appearance
    rank = np.empty_like(order)
    rank[order] = np.arange(order.shape[0])
    return rank[inv], first[order]


# ------------------------------------------------------------------------------------------------ the tiling
class Tiling:
    """vertices [Nv, 2] float64; v2g [Nv, 3] grains of every vertex (1-based, ascending); regions: per grain (1-based, dict
    order of the reference's `regions`) the ccw vertex list; centers [Ng, 2] indexed by grain - 1; edges [E, 2] directed
    joint-joint edges in the reference's order; groups: n -> (grains, ccw vertex ids [R, n], unwrapped coordinates [R, n, 2]);
    j2v_tri / j2v_vert: the reference's joint2vertex dict (sorted grain triple -> vertex) in dict order."""

    def polygons(self):
        """{grain: [n, 2] ccw polygon} in the reference's `region_coors` dict order (the draw order of plot_polygons)."""
        rows = {}
        for n, (grains, _, sc) in self.groups.items():
            for r, g in enumerate(grains.tolist()):
                rows[g] = sc[r]
        return {int(g): rows[int(g)] for g in self.region_order.tolist() if int(g) in rows}


def _voronoi_regions(lxd, seed, noise, decimals, images='all'):
    """random_voronoi_periodic (graph_datastruct.py:350-465) -> (vertices [Nv,2], (vertex, grain) incidences in the order the
    reference adds them, sorted vertex tuple of every grain).
    images='all': the reference's point set (every seed with its 8 periodic images: 9 N points through Qhull) and its region
    walk — grains and vertices are numbered exactly as the reference numbers them.
    images='margin': only the image points within 8 lattice spacings of the unit square (~N points through Qhull) and the region
    of every in-domain seed, in seed order: the same tiling (the cells of the in-domain seeds are complete), numbered differently."""
    from scipy.spatial import Voronoi
    density = 4 / lxd                                                 # ini_grain_size / lxd (:259)
    noise_eff = noise / lxd / (lxd / PATCH)                           # :260
    pts, seeds = lattice_points(density, noise_eff)
    if images == 'margin':
        m = 8 * density
        img = pts.reshape(-1, 9, 2)[:, 1:, :].reshape(-1, 2)
        near = (img[:, 0] > -m) & (img[:, 0] < 1 + m) & (img[:, 1] > -m) & (img[:, 1] < 1 + m)
        vor = Voronoi(np.concatenate([seeds, img[near]]))
        regions = [vor.regions[r] for r in vor.point_region[:seeds.shape[0]].tolist()]
    elif images == 'all':
        vor = Voronoi(pts)
        regions = vor.regions
    else:
        raise ValueError(images)
    flat, off = _ragged(regions)
    nreg = len(regions)
    lens = np.diff(off)
    seg = np.repeat(np.arange(nreg), lens)
    vx = np.where(flat >= 0, vor.vertices[np.maximum(flat, 0), 0], np.nan)
    vy = np.where(flat >= 0, vor.vertices[np.maximum(flat, 0), 1], np.nan)
    bad = (flat < 0) | (vx <= -0.5 - EPS) | (vy <= -0.5 - EPS) | (vx >= 1.5 + EPS) | (vy >= 1.5 + EPS)      # :369-377
    bad_reg = np.bincount(seg, weights=bad, minlength=nreg) > 0
    ok_reg = (~bad_reg) & (lens > 0)                                  # :383
    keep = ok_reg[seg]
    flat, seg, vx, vy = flat[keep], seg[keep], vx[keep], vy[keep]
    # vertex numbering: first appearance of the rounded coordinate pair, over ALL valid regions in order (:395-406; a
    # region that later proves to be a periodic duplicate has already registered its vertices)
    px, py = np.round(vx % 1, decimals), np.round(vy % 1, decimals)
    vid, first = _first_seen_ids(np.stack([px, py], axis=1))
    vertices = np.stack([px[first], py[first]], axis=1)
    # grains: first appearance of the SORTED vertex-id tuple (:408-414)
    reg_ids = np.nonzero(ok_reg)[0]
    rlen = lens[reg_ids]
    roff = np.zeros(reg_ids.shape[0] + 1, dtype=np.int64)
    np.cumsum(rlen, out=roff[1:])
    rseg = np.repeat(np.arange(reg_ids.shape[0]), rlen)
    maxlen = int(rlen.max())
    order = np.lexsort((vid, rseg))                                   # sort vertex ids inside every region
    pos = np.arange(vid.shape[0]) - roff[rseg]
    padded = np.full((reg_ids.shape[0], maxlen), -1, dtype=np.int64)
    padded[rseg, pos] = vid[order]
    gid, gfirst = _first_seen_ids(padded)
    # vertex2joint[v].add(alpha) for every vertex of every NEW region, in region order (:416-418)
    is_new = np.zeros(reg_ids.shape[0], dtype=bool)
    is_new[gfirst] = True
    sel = is_new[rseg]
    v_of, a_of = vid[sel], gid[rseg[sel]] + 1                         # alpha is 1-based
    return vertices, v_of, a_of, padded[gfirst]



def build_tiling(lxd, seed=1, noise=0.01, decimals=None, images=None):
    """graph.__init__ (randInit, periodic) up to and including update(init=True), without the raster."""
    if decimals is None:
        decimals = 4 + max(0, int(math.ceil(math.log10(lxd / 240.0 - 1e-9)))) if lxd > 240 else 4
    if images is None:
        images = 'all' if lxd <= 480 else 'margin'
    np.random.seed(seed)
    vertices, v_of, a_of, region_sorted = _voronoi_regions(lxd, seed, noise, decimals, images)
    nv0 = vertices.shape[0]
    # vertex2joint as Python structures only where the reference's set / dict semantics decide something: per vertex the set of
    # grains.  Vertex ids are dict keys in order of first `add` (= first appearance in a NEW region).
    order = np.argsort(v_of, kind='stable')
    vs, as_ = v_of[order], a_of[order]
    bounds = np.flatnonzero(np.diff(vs)) + 1
    starts = np.concatenate([[0], bounds])
    ends = np.concatenate([bounds, [vs.shape[0]]])
    vkeys = vs[starts]
    # dict insertion order of vertex2joint = order of each vertex's first add
    first_add = np.full(nv0, np.iinfo(np.int64).max, dtype=np.int64)
    first_add[v_of[::-1]] = np.arange(v_of.shape[0])[::-1]           # assignments in reverse: the first occurrence is written last
    key_order = vkeys[np.argsort(first_add[vkeys], kind='stable')]
    cnt = np.zeros(nv0, dtype=np.int64)
    cnt[vkeys] = ends - starts
    lo = np.zeros(nv0, dtype=np.int64)
    lo[vkeys] = starts

    def grains_of(v):                                                 # in insertion order (ascending alpha: regions come in order)
        return as_[lo[v]:lo[v] + cnt[v]]

    v2j = {}                                                          # only the vertices the quadruple pass touches
    extra_vertices = []                                               # coordinates of the vertices the pass appends
    quadruples = {}
    many = key_order[cnt[key_order] > 3]
    n_keys = int(key_order.shape[0])
    new_keys = []
    for k in many:                                                    # graph_datastruct.py:430-461, statement for statement
        k = int(k)
        v = set()
        for a in grains_of(k):                                        # the reference's set, built by the same adds
            v.add(int(a))
        grains = list(v)
        num_vertices = n_keys + len(new_keys)                         # len(self.vertex2joint)
        first = grains[0]
        v.remove(first)
        v2j[num_vertices] = v.copy()
        new_keys.append(num_vertices)
        v.add(first)
        extra_vertices.append((num_vertices, vertices[k]))
        n1 = set(int(t) for t in region_sorted[first - 1] if t >= 0)
        remove_grain = None
        for test_grains in grains[1:]:
            if len(n1.intersection(int(t) for t in region_sorted[test_grains - 1] if t >= 0)) == 1:
                remove_grain = test_grains
                break
        v.remove(remove_grain)
        v2j[k] = v.copy()
        v.remove(first)
        vv = list(v)
        quadruples.update({vv[0]: (k, num_vertices), vv[1]: (k, num_vertices)})
    # NOTE: the reference numbers an appended vertex len(vertex2joint), which equals the number of DISTINCT vertex keys so far,
    # not max id + 1; with first-seen numbering the keys are 0 .. n_keys-1 exactly when every registered vertex belongs to some
    # new region, which holds (a duplicate region's vertices are images of a new region's vertices, same rounded coordinates).
    nv = n_keys + len(new_keys)
    if n_keys != nv0:
        raise RuntimeError(f'{nv0 - n_keys} registered vertices belong to no region (reference would index out of order)')
    allv = np.zeros((nv, 2), dtype=np.float64)
    allv[:nv0] = vertices
    for vid_, xy in extra_vertices:
        allv[vid_] = xy
    # triples (sorted) per vertex, in dict order: original keys in first-add order, then the appended ones
    dict_keys = np.concatenate([key_order, np.array(new_keys, dtype=np.int64)]) if new_keys else key_order
    tri = np.zeros((nv, 3), dtype=np.int64)
    plain = key_order[cnt[key_order] == 3]
    idx = lo[plain][:, None] + np.arange(3)[None, :]
    tri[plain] = np.sort(as_[idx], axis=1)
    odd = [int(k) for k in key_order if cnt[k] != 3 and int(k) not in v2j]
    if odd:
        raise RuntimeError(f'vertices with {sorted(set(int(cnt[k]) for k in odd))} grains (not 3): the tiling is not trivalent '
                           f'at this rounding (decimals={decimals}); the reference prints them and fails later')
    for k, v in v2j.items():
        if len(v) != 3:
            raise RuntimeError(f'vertex {k} keeps {len(v)} grains after quadruple splitting')
        tri[k] = sorted(v)
    t = Tiling()
    t.lxd, t.seed, t.decimals, t.images = lxd, seed, decimals, images
    t.vertices, t.v2g, t.quadruples = allv, tri, quadruples
    # joint2vertex = dict((tuple(sorted(v)), k) for k, v in vertex2joint.items()) (:278): a repeated triple keeps its first
    # position and takes the last vertex
    keys = tri[dict_keys]
    uniq, firstpos, inv = np.unique(keys, axis=0, return_index=True, return_inverse=True)
    inv = inv.reshape(-1)
    if uniq.shape[0] != keys.shape[0]:
        lastv = np.zeros(uniq.shape[0], dtype=np.int64)
        lastv[inv] = dict_keys                                        # later assignments win, as in the dict
        o = np.argsort(firstpos, kind='stable')
        j2v_tri, j2v_vert = uniq[o], lastv[o]
    else:
        j2v_tri, j2v_vert = keys, dict_keys
    t.j2v_tri, t.j2v_vert = j2v_tri, j2v_vert
    _update_init(t)
    return t


def _set_iter_order3(tri):
    """Iteration order of Python's set(k) for the sorted triples k (graph_datastruct.py:673 `for region in set(k)`): CPython
    small-int sets of 3 elements live in an 8-slot table, slot = value & 7 with linear probing over the next 9 slots then
    perturbed probing; insertion order is the tuple order.  Evaluated by Python itself per distinct (a&7, b&7, c&7, ...) class
    would miss the perturbation, so: vectorised for the collision-free case (all three low-3-bit values distinct: order by
    value & 7), Python's own set for the rest (~1/3 of the triples)."""
    low = tri & 7
    distinct = (low[:, 0] != low[:, 1]) & (low[:, 0] != low[:, 2]) & (low[:, 1] != low[:, 2])
    out = np.empty_like(tri)
    o = np.argsort(low, axis=1, kind='stable')
    out[distinct] = np.take_along_axis(tri, o, axis=1)[distinct]
    rest = np.flatnonzero(~distinct)
    if rest.size:
        out[rest] = [list(set(r)) for r in map(tuple, tri[rest].tolist())]
    return out


def _mean_rows(a):
    """np.mean of every row the way the reference's `np.mean(x)` on a tuple of float64 sums it (add.reduce: first element, then
    numpy's pairwise sum of the rest; rows are short, so that is the contiguous inner loop of a row-wise reduction)."""
    return np.add.reduce(np.ascontiguousarray(a), axis=1) / a.shape[1]


def _update_init(t):
    """graph.update(init=True) (graph_datastruct.py:654-757) on the whole tiling at once, grouped by vertices per grain."""
    tri, vert = t.j2v_tri, t.j2v_vert
    nj = tri.shape[0]
    ng = int(tri.max())
    # regions[region].append(v) for k, v in joint2vertex.items(): for region in set(k)   (:672-677)
    it = _set_iter_order3(tri)                                        # [nj, 3] grains in set-iteration order
    g_flat = it.reshape(-1)
    v_flat = np.repeat(vert, 3)
    # dict order of `regions` = first appearance of every grain in g_flat
    firstpos = np.full(ng + 1, np.iinfo(np.int64).max, dtype=np.int64)
    firstpos[g_flat[::-1]] = np.arange(g_flat.shape[0])[::-1]         # the first occurrence is written last
    present = np.flatnonzero(firstpos[1:] < np.iinfo(np.int64).max) + 1
    region_order = present[np.argsort(firstpos[present], kind='stable')]
    o = np.argsort(g_flat, kind='stable')                             # vertices of every grain in append order
    gs, vs = g_flat[o], v_flat[o]
    deg = np.bincount(gs, minlength=ng + 1)
    start = np.zeros(ng + 2, dtype=np.int64)
    np.cumsum(deg, out=start[1:])
    centers = np.full((ng, 2), np.nan, dtype=np.float64)
    groups = {}                                                       # n -> (grains [R], ccw vertex ids [R, n], their moved coordinates [R, n, 2])
    edges_of = {}
    for n in np.unique(deg[region_order]):
        n = int(n)
        grains = region_order[deg[region_order] == n]
        if n <= 1:                                                    # :683 `if len(verts)<=1: continue`
            continue
        vid = vs[start[grains][:, None] + np.arange(n)[None, :]]      # [R, n]
        xy = t.vertices[vid]                                          # [R, n, 2]
        moved = xy.copy()
        for i in range(1, n):                                         # periodic_move(verts[i], verts[i-1]) (:691-692, :55-72)
            rel = moved[:, i] - moved[:, i - 1]
            moved[:, i] = moved[:, i] + (-1 * (rel > 0.5) + 1 * (rel < -0.5))
        inb = np.all(moved > -EPS, axis=1)                            # [R, 2] (:698-701)
        moved = moved + (1 * (~inb))[:, None, :]                      # :702-704
        c = np.stack([_mean_rows(moved[:, :, 0]), _mean_rows(moved[:, :, 1])], axis=1)       # :706-708
        centers[grains - 1] = c
        vec = moved - c[:, None, :]
        ln = np.hypot(vec[:, :, 0], vec[:, :, 1])
        ang = np.arctan2(vec[:, :, 1], vec[:, :, 0])
        ang = np.where(ang < 0, 2 * math.pi + ang, ang)
        ang = np.where(ln == 0, -math.pi, ang)                        # counterclock (:100-116)
        ln0 = np.where(ln == 0, 0.0, ln)
        srt = np.lexsort((ln0, ang), axis=1) if n > 1 else np.zeros((len(grains), 1), dtype=np.int64)
        # `sorted` is stable on the index list: lexsort is stable too
        sv = np.take_along_axis(vid, srt, axis=1)
        sc = np.take_along_axis(moved, srt[:, :, None], axis=1)
        e = np.stack([sv, np.roll(sv, -1, axis=1)], axis=2)           # [R, n, 2] cur -> nxt (:732-741)
        groups[n] = (grains, sv, sc)
        edges_of[n] = (grains, e)
    # quadruples (:736-754): a grain listed there whose ring passes through one of the two split vertices, with an edge
    # whose end points share fewer than two grains, swaps the two vertices in all its edges
    fix = {}
    if t.quadruples:
        v2g = t.v2g
        row_of = {}
        for n, (grains, sv_, _) in groups.items():
            hit = np.flatnonzero(np.isin(grains, np.fromiter(t.quadruples.keys(), dtype=np.int64)))
            for r in hit.tolist():
                row_of[int(grains[r])] = sv_[r]
        for g, (v1, v2) in t.quadruples.items():
            if g not in row_of:
                continue
            sv = row_of[g]
            save = True
            nn = len(sv)
            for i in range(nn):
                cur, nxt = int(sv[i]), int(sv[(i + 1) % nn])
                if cur in (v1, v2) or nxt in (v1, v2):
                    if len(set(v2g[cur]).intersection(set(v2g[nxt]))) != 2:
                        save = False
            if not save:
                fix[g] = (v1, v2)
    # self.edges in region dict order
    rank = np.empty(ng + 1, dtype=np.int64)
    rank[region_order] = np.arange(region_order.shape[0])
    n_of = deg[region_order]
    n_eff = np.where(n_of > 1, n_of, 0)
    eoff = np.zeros(region_order.shape[0] + 1, dtype=np.int64)
    np.cumsum(n_eff, out=eoff[1:])
    edges = np.empty((int(eoff[-1]), 2), dtype=np.int64)
    for n, (grains, e) in edges_of.items():
        at = eoff[rank[grains]][:, None] + np.arange(n)[None, :]
        edges[at.reshape(-1)] = e.reshape(-1, 2)
    for g, (v1, v2) in fix.items():
        s0 = int(eoff[rank[g]])
        blk = edges[s0:s0 + int(deg[g])]
        a, b = blk == v1, blk == v2
        blk[a], blk[b] = v2, v1
    t.n_grain, t.n_joint = ng, nj
    t.centers, t.edges, t.region_order, t.groups = centers, edges, region_order, groups
    t.deg = deg


# ------------------------------------------------------------------------------------------------ areas
def polygon_area_counts(t, imagesize):
    """Stand-in for the pixel count of the reference's raster (graph_datastruct.py:287-288 after :553-610): the exact area of
    every grain polygon in pixels of the (imagesize x imagesize) domain grid."""
    area = np.zeros(t.n_grain + 1, dtype=np.float64)
    for n, (grains, _, c) in t.groups.items():
        x, y = c[:, :, 0], c[:, :, 1]
        area[grains] = 0.5 * np.abs((x * np.roll(y, -1, axis=1)).sum(axis=1) - (y * np.roll(x, -1, axis=1)).sum(axis=1))
    return area * float(imagesize) ** 2


def _periodic_dist(p, pc, libm_pow=False):
    """periodic_dist_ (graph_datastruct.py:75-85), rows of [n, 2].  The reference squares numpy SCALARS with `**2`, which is
    libm's pow(); pow(d, 2) differs from the correctly rounded d * d by one ulp for ~1 in 1000 values.  libm_pow=True squares
    through Python floats (the same libm call; ~4 M values/s) and reproduces the reference's float64 bit for bit; the default
    vectorised product is within one float64 ulp of it and identical after the loader's float32 cast (data_loader.py:37-40)."""
    x, y, xc, yc = p[:, 0].copy(), p[:, 1].copy(), pc[:, 0], pc[:, 1]
    x = np.where(x < xc - 0.5 - EPS, x + 1, x)
    x = np.where(x > xc + 0.5 + EPS, x - 1, x)
    y = np.where(y < yc - 0.5 - EPS, y + 1, y)
    y = np.where(y > yc + 0.5 + EPS, y - 1, y)
    dx, dy = x - xc, y - yc
    if libm_pow:
        sx = np.fromiter((v ** 2 for v in dx.tolist()), dtype=np.float64, count=dx.shape[0])
        sy = np.fromiter((v ** 2 for v in dy.tolist()), dtype=np.float64, count=dy.shape[0])
        return np.sqrt(sx + sy)
    return np.sqrt(dx * dx + dy * dy)


# ------------------------------------------------------------------------------------------------ the heterograph
def generate_graph(lxd=40, seed=1, G=10.0, R=2.0, span=6, frames=120, noise=0.01, area='polygon', area_counts_fn=None,
                   decimals=None, libm_pow=None, images=None):
    """The HeteroGraph state `graph_trajectory.py --mode=generate --lxd --seed --G --R --frame` pickles, as a dict of numpy
    arrays: feature_dicts {'grain' [Ng, 11], 'joint' [Nj, 8]} (float64), edge_index_dicts (int64 [2, E] per edge type),
    edge_weight_dicts ([E, 1]), mask, plus 'tiling' (the Tiling) and 'span'.  `span` is the reference's nearest-neighbour lookup
    in GR_train_grid.pkl (graph_trajectory.py:1308-1316; 6 for G = 10, R = 2): pass it, or use `span_from_grid`.
    area: 'polygon' | 'raster' (area_counts_fn(tiling, imagesize) -> {grain: pixel count}).
    libm_pow: see _periodic_dist; None = on up to lxd 240 (the sizes the reference itself reaches).
    images: see _voronoi_regions; None = 'all' (the reference's numbering) up to lxd 480, 'margin' above."""
    if libm_pow is None:
        libm_pow = lxd <= 240
    t = build_tiling(lxd, seed, noise, decimals, images)
    ng, nj = t.n_grain, t.n_joint
    # orientations (graph_datastruct.py:292-305) — drawn after update(init) from the same stream
    ux, uy, uz = np.random.randn(ng), np.random.randn(ng), np.random.randn(ng)
    theta_x = np.arctan2(uy, ux) % (math.pi / 2)
    theta_z = np.arctan2(np.sqrt(ux ** 2 + uy ** 2), uz) % (math.pi / 2)
    imagesize = int(lxd / MESH) + 1
    if area == 'raster':
        if area_counts_fn is None:
            raise ValueError("area='raster' needs area_counts_fn(tiling, imagesize) -> {grain: pixels}")
        counts = area_counts_fn(t, imagesize)
    elif area == 'polygon':
        counts = polygon_area_counts(t, imagesize)
    else:
        raise ValueError(area)
    s = int(np.round(PATCH / MESH)) + 1                               # graph_trajectory.py:909
    grain = np.zeros((ng, 10))
    joint = np.zeros((t.vertices.shape[0], 6))
    gmask = np.zeros((ng, 1), dtype=int)
    jmask = np.zeros((t.vertices.shape[0], 1), dtype=int)
    have = ~np.isnan(t.centers[:, 0])
    grain[have, 0:2] = t.centers[have]
    if isinstance(counts, dict):
        cnt = np.zeros(ng)
        for g, c in counts.items():
            if 1 <= g <= ng:
                cnt[g - 1] = c
    else:                                                             # array indexed by grain id (1-based)
        cnt = np.asarray(counts, dtype=np.float64)[1:ng + 1]
    grain[have, 3] = cnt[have] / s ** 2
    gmask[have, 0] = 1
    grain[:, 2] = 0 / frames
    grain[:, 5], grain[:, 6] = np.cos(theta_x), np.sin(theta_x)
    grain[:, 7], grain[:, 8] = np.cos(theta_z), np.sin(theta_z)
    joint[:, 0:2] = t.vertices
    jmask[:, 0] = 1
    joint[:, 2] = 0 / frames
    joint[:, 3] = 1 - G / 10
    joint[:, 4] = R / 2
    # edges (graph_trajectory.py:958-979)
    gj_g = (t.j2v_tri - 1).reshape(-1)
    gj_j = np.repeat(t.j2v_vert, 3)
    gj = np.stack([gj_g, gj_j])
    gj_len = _periodic_dist(t.vertices[gj_j], t.centers[gj_g], libm_pow)
    jj = t.edges.T.copy()
    jj_len = _periodic_dist(t.vertices[jj[0]], t.vertices[jj[1]], libm_pow)
    # form_gradient(None, None, ...) + append_history([]) (graph_datastruct.py:978-1011)
    grain[:, 4] *= 20
    grain[:, 9] = span / 120
    joint[:, 5] = span / 120
    grain = np.hstack((grain, 0 * grain[:, :1]))
    joint = np.hstack((joint, 0 * joint[:, :2]))
    return {'feature_dicts': {'grain': grain, 'joint': joint},
            'edge_index_dicts': {ET_GJ: gj, ET_JG: gj[::-1].copy(), ET_JJ: jj},
            'edge_weight_dicts': {ET_GJ: gj_len[:, None], ET_JG: gj_len[:, None], ET_JJ: jj_len[:, None]},
            'mask': {'grain': gmask, 'joint': jmask}, 'span': span, 'tiling': t,
            'physical_params': {'G': G, 'R': R, 'seed': seed, 'height': 0}}


def span_from_grid(G, R, grid):
    """graph_trajectory.py:1314-1316: nearest (G, R) of the training grid.  grid = the dict pickled in GR_train_grid.pkl."""
    from scipy.interpolate import griddata
    G_ = (G - grid['G_min']) / (grid['G_max'] - grid['G_min'])
    R_ = (R - grid['R_min']) / (grid['R_max'] - grid['R_min'])
    return griddata(np.array([grid['G'], grid['R']]).T, np.array(grid['span']), (G_, R_), method='nearest')


def model_inputs(hg, lxd):
    """What the rollout driver feeds the models at the first step: the loader's tensors (data_loader.py:113-162: float32 /
    int64) after the patch scaling of test.py:29-55 when the domain is larger than one 40-um patch (:310-312).
    -> (x_dict, edge_index_dict, edge_attr_dict, geometry) with geometry = {'domain_factor', 'domain_offset' [Nj,2],
    'grain_coor_offset' [Ng,2], 'global' {type: [N,2] global position in [0,1)^2}}."""
    import torch
    x = {k: torch.FloatTensor(v) for k, v in hg['feature_dicts'].items()}
    ei = {k: torch.LongTensor(np.ascontiguousarray(v)) for k, v in hg['edge_index_dicts'].items()}
    ea = {k: torch.FloatTensor(v) for k, v in hg['edge_weight_dicts'].items()}
    glob = {k: v[:, :2].clone() for k, v in x.items()}
    factor = lxd / PATCH
    geom = {'domain_factor': factor, 'domain_offset': 0, 'global': glob}
    if factor > 1:
        for k in ea:
            ea[k] *= factor
        x['grain'][:, :2] *= factor
        x['joint'][:, :2] *= factor
        off = torch.floor(x['joint'][:, :2])
        x['joint'][:, :2] = x['joint'][:, :2] - off
        goff = x['grain'][:, :2] - x['grain'][:, :2] % 1
        x['grain'][:, :2] = x['grain'][:, :2] - goff
        geom.update({'domain_offset': off, 'grain_coor_offset': goff})
    return x, ei, ea, geom
