"""Row f4: the polygon raster (graph_datastruct.py:553-610) and the layer error (:346-348).

The reference's raster is PIL's; oracle/raster_oracle.plot_polygons calls PIL the way the reference does (and equals the
reference's alpha_field, tests/test_generate.py).  Here: the scan-line rule the device kernel implements, restated in Python, is
pinned against PIL (CPU), and the kernel is compared with PIL and with the restated rule (GPU)."""
import os

import numpy as np
import pytest
import torch

import raster_oracle as ro
from util import GOLDEN

from graingraphnn_b200 import generate as G


def _convex(rng, lo=-6, hi=70):
    from scipy.spatial import ConvexHull
    while True:
        P = rng.integers(lo, hi, size=(int(rng.integers(3, 10)), 2))
        try:
            h = ConvexHull(P)
        except Exception:
            continue
        pts = [tuple(int(v) for v in P[i]) for i in h.vertices]
        return pts[::-1] if rng.random() < 0.5 else pts


def _pil(W, H, pts):
    from PIL import Image, ImageDraw
    im = Image.new('L', (W, H))
    ImageDraw.Draw(im).polygon(pts, fill=1)
    return np.array(im)


@pytest.mark.parametrize('lxd,seed', [(40, 1), (40, 10020), (120, 0)])
def test_restated_scanline_rule_equals_pil_on_the_reference_tilings(lxd, seed):
    t = G.build_tiling(lxd, seed)
    s = int(lxd / 0.08) + 1
    polys = t.polygons()
    np.testing.assert_array_equal(ro.plot_polygons_restated(polys, s), ro.plot_polygons(polys, s))


def test_restated_scanline_rule_on_random_convex_polygons():
    rng = np.random.default_rng(1)
    bad = 0
    for _ in range(1500):
        pts = _convex(rng)
        img = np.zeros((64, 64), dtype=np.uint8)
        ro.fill_polygon_restated(img, pts, 1)
        bad += not np.array_equal(img, _pil(64, 64, pts))
    assert bad <= 0.02 * 1500, bad          # PIL has an order-dependent corner rule on isolated polygons; tilings overdraw it


@pytest.mark.gpu
@pytest.mark.parametrize('lxd,seed', [(40, 1), (120, 0), (240, 1)])
def test_device_raster_equals_pil_on_the_reference_tilings(lxd, seed):
    from graingraphnn_b200 import raster
    t = G.build_tiling(lxd, seed)
    s = int(lxd / 0.08) + 1
    polys = t.polygons()
    alpha = raster.plot_polygons(polys, s, 'cuda:0')
    ref = ro.plot_polygons(polys, s)
    assert np.array_equal(alpha.cpu().numpy(), ref)
    if lxd == 40:
        z = np.load(os.path.join(GOLDEN, 'generate_lxd40.npz'))
        assert np.array_equal(alpha.cpu().numpy(), z['alpha_field'])             # the reference's own alpha_field
        assert raster.area_counts(alpha) == dict(zip(z['area_ids'].tolist(), z['area_counts'].tolist()))
    # layer error against a shifted copy, as compute_error_layer counts it
    other = torch.roll(alpha, 3, dims=1)
    assert raster.error_layer(other, alpha) == float(np.sum(np.roll(ref, 3, axis=1) != ref) / ref.size)


@pytest.mark.gpu
def test_device_raster_equals_the_restated_rule_on_random_polygons():
    from graingraphnn_b200 import raster
    rng = np.random.default_rng(7)
    s = 32                                                                        # image 64 x 64, polygons in [-6, 70)
    for _ in range(200):
        pts = _convex(rng)
        poly = {5: (np.array(pts, dtype=np.float64) + 0.25) / s}                  # int((v + 0.25) / s * s) == v
        got = raster.plot_polygons(poly, s, 'cuda:0').cpu().numpy()
        np.testing.assert_array_equal(got, ro.plot_polygons_restated(poly, s))
