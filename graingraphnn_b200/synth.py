"""Scalable synthetic grain graphs for the 10^5-10^6-grain configurations.

`honeycomb_graph` builds, in O(N) vectorised numpy, the graph a periodic hexagonal seed lattice induces — the same
lattice `graph_trajectory.py --mode=generate` starts from (graph_datastruct.py:118-160: triangular lattice of spacing
4 um = 0.1 patch, Gaussian jitter) — with joints at the circumcentres of the jittered Delaunay triangles (= the Voronoi
vertices for as long as the jitter does not flip a Delaunay edge).  Every joint has exactly 3 joint and 3 grain
neighbours, every grain 6 joints: Nj = 2 Ng, E = 6 Ng per edge type, like the reference's fixtures.  Features follow
`form_states_tensor(0)` (graph_trajectory.py:901-955): coordinates in 40-um patch units wrapped mod 1 (test.py:29-55),
z = 0, area in patch units, random orientations, G' = 1 - G/10, R' = R/2, span = 6/120.
The reference generator itself is O(N^2) with a (lxd/0.08)^2 raster and cannot reach these sizes.
"""
import math

import numpy as np
import torch

ET_GJ, ET_JG, ET_JJ = ('grain', 'push', 'joint'), ('joint', 'pull', 'grain'), ('joint', 'connect', 'joint')


def lattice_dims(patches_x, patches_y, spacing=0.1):
    """Grain columns / rows of a triangular lattice of ~`spacing` patches covering patches_x x patches_y patches."""
    nx = int(round(patches_x / spacing))
    ny = int(round(patches_y / (spacing * math.sqrt(3) / 2)))
    ny += ny % 2
    return nx, ny


def honeycomb_graph(nx, ny, seed=0, jitter=0.08, G=10.0, R=2.0, span=6, patches=None, return_global=False):
    """nx x ny grains (ny even) on a torus.  Returns (x_dict, edge_index_dict) as CPU torch tensors
    (x float32, edge_index int64 [2, 6*nx*ny]); node order is row-major in space.
    patches = (px, py): domain size in patch units (default: spacing 0.1 patch in x, 0.1*sqrt(3)/2 in y, rounded up to
    whole patches so that the periodic wrap coincides with a patch wrap)."""
    assert ny % 2 == 0 and nx >= 2 and ny >= 2
    rng = np.random.default_rng(seed)
    n = nx * ny
    if patches is None:
        patches = (max(1, math.ceil(nx * 0.1 - 1e-9)), max(1, math.ceil(ny * 0.1 * math.sqrt(3) / 2 - 1e-9)))
    Lx, Ly = float(patches[0]), float(patches[1])
    ax, ay = Lx / nx, Ly / ny
    jj_, ii_ = np.meshgrid(np.arange(ny), np.arange(nx), indexing='ij')       # [ny, nx]
    i, j = ii_.ravel(), jj_.ravel()
    o = j & 1
    sx = (i + 0.5 * o) * ax + rng.normal(0.0, jitter * ax, n)
    sy = j * ay + rng.normal(0.0, jitter * ax, n)

    def sid(ii, jj2):
        return (jj2 % ny) * nx + (ii % nx)

    s0 = sid(i, j)
    right = sid(i + 1, j)
    up = sid(i + o, j + 1)
    down = sid(i + o, j - 1)
    tri = np.empty((2 * n, 3), dtype=np.int64)            # joint 2s = T_A(s) = {s, right, up}; 2s+1 = T_B(s) = {s, right, down}
    tri[0::2] = np.stack([s0, right, up], 1)
    tri[1::2] = np.stack([s0, right, down], 1)

    def unwrap(d, L):
        return d - L * np.round(d / L)

    # circumcentres relative to the first site of each triangle
    px, py = sx[tri[:, 0]], sy[tri[:, 0]]
    bx, by = unwrap(sx[tri[:, 1]] - px, Lx), unwrap(sy[tri[:, 1]] - py, Ly)
    cx, cy = unwrap(sx[tri[:, 2]] - px, Lx), unwrap(sy[tri[:, 2]] - py, Ly)
    d = 2.0 * (bx * cy - by * cx)
    ux = (cy * (bx * bx + by * by) - by * (cx * cx + cy * cy)) / d
    uy = (bx * (cx * cx + cy * cy) - cx * (bx * bx + by * by)) / d
    jx, jy = (px + ux) % Lx, (py + uy) % Ly

    # joint-joint edges: T_A(i,j) touches T_B(i,j), T_B(i-1+o,j+1), T_B(i+o,j+1); both directions
    A = 2 * s0
    nb = np.stack([2 * s0 + 1, 2 * sid(i - 1 + o, j + 1) + 1, 2 * sid(i + o, j + 1) + 1], 1)     # [n,3]
    src = np.concatenate([np.repeat(A, 3), nb.ravel()])
    dst = np.concatenate([nb.ravel(), np.repeat(A, 3)])
    order = np.lexsort((src, dst))                         # sorted by target then source, like the fixtures (by joint)
    ei_jj = np.stack([src[order], dst[order]])
    # grain-joint edges: per joint, its 3 grains (graph_trajectory.py:958-962 ordering: grouped by joint)
    joint_ids = np.repeat(np.arange(2 * n), 3)
    ei_gj = np.stack([tri.ravel(), joint_ids])
    ei_jg = ei_gj[::-1].copy()

    # grains: hexagon = circumcentres of the 6 incident triangles, counter-clockwise
    hexj = np.stack([2 * s0, 2 * sid(i - 1 + o, j + 1) + 1, 2 * sid(i - 1, j), 2 * sid(i - 1, j) + 1,
                     2 * sid(i - 1 + o, j - 1), 2 * s0 + 1], 1)                                  # [n,6]
    hx = unwrap(jx[hexj] - sx[:, None], Lx)
    hy = unwrap(jy[hexj] - sy[:, None], Ly)
    area = 0.5 * np.abs((hx * np.roll(hy, -1, 1) - np.roll(hx, -1, 1) * hy).sum(1))
    gx, gy = (sx + hx.mean(1)) % Lx, (sy + hy.mean(1)) % Ly

    ux_, uy_, uz_ = rng.standard_normal(n), rng.standard_normal(n), rng.standard_normal(n)
    theta_x = np.arctan2(uy_, ux_) % (math.pi / 2)                       # graph_datastruct.py:292-305
    theta_z = np.arctan2(np.sqrt(ux_ ** 2 + uy_ ** 2), uz_) % (math.pi / 2)

    xg = np.zeros((n, 11), dtype=np.float32)
    xg[:, 0], xg[:, 1] = gx % 1.0, gy % 1.0
    xg[:, 3] = area
    xg[:, 5], xg[:, 6], xg[:, 7], xg[:, 8] = np.cos(theta_x), np.sin(theta_x), np.cos(theta_z), np.sin(theta_z)
    xg[:, 9] = span / 120.0
    xj = np.zeros((2 * n, 8), dtype=np.float32)
    xj[:, 0], xj[:, 1] = jx % 1.0, jy % 1.0
    xj[:, 3], xj[:, 4], xj[:, 5] = 1.0 - G / 10.0, R / 2.0, span / 120.0
    # keep wrapped coordinates strictly inside [0, 1) after the float32 cast
    for arr in (xg, xj):
        np.clip(arr[:, :2], 0.0, np.nextafter(np.float32(1.0), np.float32(0.0)), out=arr[:, :2])

    x = {'grain': torch.from_numpy(xg), 'joint': torch.from_numpy(xj)}
    ei = {ET_GJ: torch.from_numpy(np.ascontiguousarray(ei_gj)), ET_JG: torch.from_numpy(np.ascontiguousarray(ei_jg)),
          ET_JJ: torch.from_numpy(np.ascontiguousarray(ei_jj))}
    if return_global:
        glob = {'grain': torch.from_numpy(np.stack([gx / Lx, gy / Ly], 1)), 'joint': torch.from_numpy(np.stack([jx / Lx, jy / Ly], 1)),
                'patches': (Lx, Ly)}
        return x, ei, glob
    return x, ei


def lattice_domain(patches=(36, 30), seed=1):
    """(x, ei, global positions, patches) of the honeycomb stand-in covering `patches` 40-um patches — a quick regular test
    domain for side scripts; benchmarks and parity tests use generate.generate_graph (the reference's generate mode)."""
    nx, ny = lattice_dims(*patches)
    x, ei, glob = honeycomb_graph(nx, ny, seed=seed, patches=patches, return_global=True)
    return x, ei, glob, patches
