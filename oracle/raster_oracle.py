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
