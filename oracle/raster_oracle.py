"""CPU restatement of the reference's polygon raster and layer-error QoI — TEST INFRASTRUCTURE ONLY.

Follows graph_datastruct.py:553-610 (`graph.plot_polygons`, periodic branch) and :346-348 (`compute_error_layer`) on the
same third-party routine the reference calls (PIL `ImageDraw.polygon`): polygons are drawn in `region_coors` dict order
into a (2s x 2s) image with the grain id as colour, integer vertex coordinates `int(coor * s)` (truncation), and the four
s x s quadrants are folded with `max`.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline may import this.
"""
import numpy as np


def plot_polygons(polygons, s):
    """polygons: {grain id (1-based): [[x, y], ...]} in draw order; s = imagesize[0].  -> alpha_field [s, s] int (Image
    convention [ny, nx])."""
    import PIL.Image as Image
    import PIL.ImageDraw as ImageDraw
    image = Image.new('RGB', (2 * s, 2 * s))                                  # :566
    draw = ImageDraw.Draw(image)
    for region_id, poly in polygons.items():                                  # :570-589
        Rid = region_id // (255 * 255)
        Gid = (region_id - Rid * 255 * 255) // 255
        Bid = region_id - Rid * 255 * 255 - Gid * 255
        p = [tuple(np.asarray(np.array(v) * s, dtype=int)) for v in poly]
        if len(p) > 1:
            draw.polygon(p, fill=(Rid, Gid, Bid))
    img = np.array(image, dtype=int)                                          # :591-597
    img = img[:, :, 0] * 255 * 255 + img[:, :, 1] * 255 + img[:, :, 2]
    img = np.stack([img[:s, :s], img[s:, :s], img[:s, s:], img[s:, s:]])
    return np.max(img, axis=0)


def area_counts(alpha_field):
    """graph_datastruct.py:287-288 / graph_trajectory.py:1301-1302."""
    cur, counts = np.unique(alpha_field, return_counts=True)
    return dict(zip(cur.tolist(), counts.tolist()))


def error_layer(alpha_pde, alpha_field):
    """compute_error_layer (:346-348)."""
    return np.sum(alpha_pde != alpha_field) / len(alpha_pde.flatten())


# ---------------------------------------------------------------------------------------------------------------------
# The scan-line rule gg_raster_polygons implements, in plain Python with float32 intersections — pinned against PIL itself in
# tests/test_raster.py (pixel-exact on the reference's tilings) so that the device kernel can be checked against it anywhere.
def _round_up(f):
    import math
    f = float(f)
    return int(math.floor(f + 0.5)) if f >= 0 else -int(math.floor(abs(f) + 0.5))


def _round_down(f):
    import math
    f = float(f)
    return int(math.ceil(f - 0.5)) if f >= 0 else -int(math.ceil(abs(f) - 0.5))


def fill_polygon_restated(img, pts, ink):
    """Overdraw `ink` over the pixels PIL's ImageDraw.polygon(pts, fill=ink) covers (integer vertices), into img [H, W]."""
    f32 = np.float32
    H, W = img.shape
    n = len(pts)
    E = []
    ymin_p, ymax_p = min(p[1] for p in pts), max(p[1] for p in pts)
    for i in range(n):
        (x0, y0), (x1, y1) = pts[i], pts[(i + 1) % n]
        if y0 == y1:
            if 0 <= y0 < H:
                xa, xb = max(min(x0, x1), 0), min(max(x0, x1), W - 1)
                if xa <= xb:
                    img[y0, xa:xb + 1] = ink
            continue
        E.append((x0, y0, min(y0, y1), max(y0, y1), f32(f32(x1 - x0) / f32(y1 - y0))))
    for y in range(max(ymin_p, 0), min(ymax_p, H) + 1):
        xx = []
        for (x0, y0, ya, yb, dx) in E:
            if ya <= y <= yb:
                xv = f32(f32(y - y0) * dx + f32(x0))
                xx.append(xv)
                if y == yb and y < ymax_p:
                    xx.append(xv)
        if len(xx) == 2 and xx[0] == xx[1] and (y == ymax_p or y == ymin_p):
            off = -1 if y == ymax_p else 1
            adj = [f32(f32(y + off - e[1]) * e[4] + f32(e[0])) for e in E if e[2] <= y <= e[3]]
            if all(a > xx[0] for a in adj):
                xx[1] = f32(max(float(xx[0]), _round_up(min(adj)) - 1))
        xx.sort()
        x_pos = -1 if not xx else 0
        for i in range(1, len(xx), 2):
            x_end = _round_down(xx[i])
            if x_end < x_pos:
                continue
            x_start = _round_up(xx[i - 1])
            if x_pos > x_start:
                x_start = x_pos
                if x_end < x_start:
                    continue
            if 0 <= y < H:
                xa, xb = max(x_start, 0), min(x_end, W - 1)
                if xa <= xb:
                    img[y, xa:xb + 1] = ink
            x_pos = x_end + 1


def plot_polygons_restated(polygons, s):
    img = np.zeros((2 * s, 2 * s), dtype=np.int64)
    for gid, poly in polygons.items():
        p = [tuple(int(v) for v in np.asarray(np.array(c) * s, dtype=int)) for c in poly]
        if len(p) > 1:
            fill_polygon_restated(img, p, gid)
    return np.max(np.stack([img[:s, :s], img[s:, :s], img[:s, s:], img[s:, s:]]), axis=0)
