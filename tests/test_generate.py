"""Row f3: the vectorised generator against the reference's own `--mode=generate` (graph_trajectory.py:1289-1333).

Committed vectors (tests/golden/generate_lxd*.npz, made by oracle/make_golden_generate.py from the UNMODIFIED reference)
pin the sizes the reference reaches; with /root/reference present further seeds run live.  The raster the `area`
feature needs is the reference's own PIL routine restated in oracle/raster_oracle.py."""
import os

import numpy as np
import pytest

from util import ET, GOLDEN

from graingraphnn_b200 import generate as G
import raster_oracle as ro

HAVE_REF = os.path.exists('/root/reference/graph_trajectory.py')


def _raster_counts(t, s):
    c = ro.area_counts(ro.plot_polygons(t.polygons(), s))
    return {int(k): int(v) for k, v in c.items()}


def _assert_equal(hg, z):
    np.testing.assert_array_equal(hg['feature_dicts']['grain'], z['x_grain'])
    np.testing.assert_array_equal(hg['feature_dicts']['joint'], z['x_joint'])
    np.testing.assert_array_equal(hg['mask']['grain'], z['mask_grain'])
    np.testing.assert_array_equal(hg['mask']['joint'], z['mask_joint'])
    for i, e in enumerate(ET):
        np.testing.assert_array_equal(hg['edge_index_dicts'][e], z[f'ei{i}'].astype(np.int64))     # position for position
        np.testing.assert_array_equal(hg['edge_weight_dicts'][e], z[f'ew{i}'])                     # float64, bit for bit


@pytest.mark.parametrize('lxd', [40, 120, 240])
def test_generator_equals_reference_generate_mode(lxd):
    z = np.load(os.path.join(GOLDEN, f'generate_lxd{lxd}.npz'))
    hg = G.generate_graph(lxd=lxd, seed=int(z['seed']), G=10.0, R=2.0, span=int(z['span']), area='raster',
                          area_counts_fn=_raster_counts)
    _assert_equal(hg, z)
    t = hg['tiling']
    assert (t.n_grain, t.n_joint) == (z['x_grain'].shape[0], z['x_joint'].shape[0])
    # SURVEY §8c KAT (3): every joint has in-degree 3 in jj and gj; E = 3 Nj = 6 Ng
    for e in (ET[0], ET[2]):
        assert (np.bincount(hg['edge_index_dicts'][e][1], minlength=t.n_joint) == 3).all()
    assert hg['edge_index_dicts'][ET[2]].shape[1] == 3 * t.n_joint == 6 * t.n_grain
    if lxd == 120:
        assert (t.n_grain, t.n_joint) == (1043, 2086)                 # SURVEY §8c KAT (5)


def test_raster_oracle_equals_reference_alpha_field():
    z = np.load(os.path.join(GOLDEN, 'generate_lxd40.npz'))
    t = G.build_tiling(40, int(z['seed']))
    af = ro.plot_polygons(t.polygons(), 501)
    np.testing.assert_array_equal(af, z['alpha_field'])
    c = ro.area_counts(af)
    assert c == dict(zip(z['area_ids'].tolist(), z['area_counts'].tolist()))


def test_vectorised_edge_length_is_within_one_ulp_and_equal_in_float32():
    z = np.load(os.path.join(GOLDEN, 'generate_lxd240.npz'))
    hg = G.generate_graph(lxd=240, seed=int(z['seed']), libm_pow=False)          # the scalable arithmetic (d * d instead of libm pow)
    for i, e in enumerate(ET):
        a, b = hg['edge_weight_dicts'][e], z[f'ew{i}']
        assert np.all(np.abs(a - b) <= np.spacing(b))
        np.testing.assert_array_equal(a.astype(np.float32), b.astype(np.float32))


def test_polygon_area_tracks_the_pixel_count():
    z = np.load(os.path.join(GOLDEN, 'generate_lxd120.npz'))
    hg = G.generate_graph(lxd=120, seed=int(z['seed']))                           # area='polygon'
    a, b = hg['feature_dicts']['grain'][:, 3], z['x_grain'][:, 3]
    assert abs(a.sum() / b.sum() - 1) < 1e-6                                      # both tile the domain
    assert np.abs(a - b).max() < 0.06 * b.mean() and np.abs(a - b).mean() < 0.02 * b.mean()   # the raster's boundary pixels


def test_margin_images_give_the_same_tiling_up_to_numbering():
    a, b = G.build_tiling(240, 1, images='all'), G.build_tiling(240, 1, images='margin')
    assert (a.n_grain, a.n_joint) == (b.n_grain, b.n_joint)
    assert set(map(tuple, a.vertices.tolist())) == set(map(tuple, b.vertices.tolist()))

    def cells(t):
        return {frozenset(map(tuple, t.vertices[sv[r]].tolist())) for _, (g, sv, _) in t.groups.items() for r in range(len(g))}

    def jj(t):
        return {(tuple(t.vertices[i]), tuple(t.vertices[j])) for i, j in t.edges.tolist()}
    assert cells(a) == cells(b) and jj(a) == jj(b)


def test_large_domain_is_a_trivalent_tiling_with_the_reference_degree_spread():
    hg = G.generate_graph(lxd=600, seed=3)                                        # 26 k grains, margin images, 5 decimals
    t = hg['tiling']
    assert t.images == 'margin' and t.decimals == 5
    gj, jg, jj = (hg['edge_index_dicts'][e] for e in ET)
    assert (np.bincount(jj[1], minlength=t.n_joint) == 3).all() and (np.bincount(gj[1], minlength=t.n_joint) == 3).all()
    assert t.n_joint == 2 * t.n_grain and jj.shape[1] == 6 * t.n_grain            # Euler on the torus
    deg = np.bincount(jg[1], minlength=t.n_grain)
    assert deg.min() >= 3 and deg.max() <= 10 and 0.8 < (deg == 6).mean() < 0.9   # {6: 84 %, 5: 8 %, 7: 8 %}
    # jj edges come in both directions
    fwd = set(map(tuple, jj.T.tolist()))
    assert all((b, a) in fwd for a, b in list(fwd)[:2000])
    x, ei, ea, geom = G.model_inputs(hg, 600)
    assert geom['domain_factor'] == 15 and float(x['joint'][:, :2].min()) >= 0 and float(x['joint'][:, :2].max()) < 1
    assert float(ea[ET[2]].max()) < 0.5                                           # edge lengths in patch units


@pytest.mark.skipif(not HAVE_REF, reason='/root/reference is not mounted')
@pytest.mark.parametrize('lxd,seed', [(40, 10020), (80, 3), (160, 7)])
def test_generator_equals_live_reference(lxd, seed):
    from make_golden_generate import reference_generate
    traj, ref = reference_generate(lxd, seed)
    counts = {int(k): int(v) for k, v in traj.area_counts.items()}
    hg = G.generate_graph(lxd=lxd, seed=seed, span=int(ref.span), area='raster', area_counts_fn=lambda t, s: counts)
    z = {'x_grain': ref.feature_dicts['grain'], 'x_joint': ref.feature_dicts['joint'],
         'mask_grain': ref.mask['grain'], 'mask_joint': ref.mask['joint']}
    for i, e in enumerate(ET):
        z[f'ei{i}'], z[f'ew{i}'] = ref.edge_index_dicts[e], ref.edge_weight_dicts[e]
    _assert_equal(hg, z)
    assert len(hg['tiling'].quadruples) == len(traj.quadruples)
