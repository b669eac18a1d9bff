"""The rollout loop (test.py:353-577) on the resident engine — graingraphnn_b200/rollout.py.

CPU: the polygon / centre bookkeeping of the QoI path against the reference's own GNN_update + graph.update (live, when
/root/reference is mounted) and the layer-error KAT of SURVEY §8c (4).  GPU: the driver against the same loop assembled
from the oracle's pieces (NN step, event candidates, topology update, region centres, edge lengths)."""
import contextlib
import io
import os

import numpy as np
import pytest
import torch

import grain_oracle as orc
import raster_oracle as ro
import topology_oracle as topo
from util import ET, load_graph, rel_err

HAVE_REF = os.path.exists('/root/reference/graph_trajectory.py')


@pytest.mark.skipif(not HAVE_REF, reason='/root/reference is not mounted')
def test_region_polygons_equal_the_reference_update_and_layer_error_kat():
    """GNN_update(0, ...) of the 40x40 fixture trajectory: same polygons in the same draw order, same alpha_field through the
    reference's raster, error_layer(t = 0) = 0.0240 (SURVEY §8c KAT 4)."""
    import gzip
    import sys
    import dill
    import ref_shims
    from graingraphnn_b200.rollout import region_polygons
    ref_shims.install_plot_stubs(); ref_shims.add_reference_to_path()
    with contextlib.redirect_stdout(io.StringIO()):
        import graph_trajectory as gt
        import __main__
        __main__.graph_trajectory = gt.graph_trajectory
        with gzip.open('/root/reference/graphs/40_40/traj10020.pkl.gz', 'rb') as f:
            traj = dill.load(f)
        with open('/root/reference/graphs/40_40/seed10020_G1.904_R0.558_span6.pkl', 'rb') as f:
            g = dill.load(f)[0]
        x = {k: torch.FloatTensor(v) for k, v in g.feature_dicts.items()}
        ei = {k: torch.LongTensor(v) for k, v in g.edge_index_dicts.items()}
        mask = {k: torch.from_numpy(np.ones_like(v)) for k, v in g.mask.items()}
        traj.extraV_traj = []
        traj.GNN_update(0, {k: v.clone() for k, v in x.items()}, mask, True, ei, True)
        traj.plot_polygons()
    polys, centers = region_polygons(x['joint'][:, :2].numpy(), ei[ET[0]].numpy())
    assert list(polys.keys()) == list(traj.region_coors.keys())
    for gid, p in polys.items():
        np.testing.assert_allclose(p, np.array(traj.region_coors[gid], dtype=np.float64), rtol=0, atol=1e-6)
    s = traj.imagesize[0]
    alpha = ro.plot_polygons(polys, s)
    np.testing.assert_array_equal(alpha, traj.alpha_field)
    err = ro.error_layer(traj.alpha_pde, alpha)
    assert abs(err - traj.error_layer) < 1e-12 and abs(err - 0.0240) < 5e-4
    # the QoI bookkeeping of the same call (graph_trajectory.py:1041-1051, :1100-1103) against the oracle's restatement
    ac, extra, va = orc.area_bookkeeping(x['grain'], mask['grain'], ei[ET[0]], traj.lxd)
    assert ac == traj.area_traj[-1]
    np.testing.assert_array_equal(extra, traj.extraV_traj[-1])
    assert sorted(va) == sorted(int(k) for k in traj.vertex_area)              # three terms per joint, summed in another grain order:
    np.testing.assert_allclose([va[k] for k in sorted(va)], [traj.vertex_area[k] for k in sorted(traj.vertex_area)], rtol=1e-14)   # an ulp


def _oracle_loop(sd_r, sd_c, x, ei, ea, mask, steps, span, edge_thr, area_thr):
    """The frame loop of test.py:353-577 from the oracle's pieces (single patch: domain_factor 1)."""
    x = {k: v.clone() for k, v in x.items()}
    ei = {k: v.clone() for k, v in ei.items()}
    ea = {k: v.clone() for k, v in ea.items()}
    mask = {k: v.clone() for k, v in mask.items()}
    events, log = [], []
    for _ in range(steps):
        y = orc.regressor_forward(sd_r, x, ei, ea)
        y.update(orc.classifier_forward(sd_c, x, ei, ea))
        orc.regressor_update(x, y, span)
        L1, ge = orc.event_candidates(y, ei[ET[2]], mask['grain'], edge_thr, area_thr)
        y['grain_event'] = ge
        pairs = torch.zeros(0, 2, dtype=torch.int64)
        if len(L1) or len(ge):
            act_g = (y['grain'][:, 0] > -10).nonzero().view(-1)
            act_j = (y['joint'][:, 0] > -10).nonzero().view(-1)
            _, new_ei, pairs = topo.topology_update(x, ei, y, mask, act_g, act_j, threshold=edge_thr)
            if len(y['grain_event']) or len(pairs):
                ei = new_ei
        events.extend(int(g) for g in y['grain_event'])
        cen = orc.region_center(x['joint'], ei[ET[0]], x['grain'].shape[0])
        orc.grain_xy_writeback(x['grain'], cen)
        ea = orc.edge_attr_rebuild(x, ei)
        log.append((len(pairs), {k: v.clone() for k, v in x.items()}, {k: v.clone() for k, v in ei.items()}))
    return events, log, x, ei, mask


@pytest.mark.gpu
@pytest.mark.parametrize('rows', ['caller', 'morton'])
def test_rollout_driver_matches_the_reference_order_loop(rows):
    from graingraphnn_b200.engine import RolloutEngine
    from graingraphnn_b200.rollout import RolloutDriver
    x, ei, ea = load_graph('c1')
    sd_r, sd_c = orc.synth_state_dict('regressor', 1), orc.synth_state_dict('classifier', 2)
    sd_c['lin2.bias'] = sd_c['lin2.bias'] - 0.613           # the seeded logits sit in 1.00 .. 1.02: move the 0.6 threshold into their upper tail
    mask = {'grain': torch.ones(x['grain'].shape[0], 1), 'joint': torch.ones(x['joint'].shape[0], 1)}
    edge_thr, area_thr, span, steps = 0.6, 0.0235, 6, 4     # area threshold: 3 of the 118 grains fall below it in 4 steps
    ev_ref, log, x_ref, ei_ref, mask_ref = _oracle_loop(sd_r, sd_c, x, ei, ea, mask, steps, span, edge_thr, area_thr)
    assert sum(n for n, _, _ in log) > 0 and len(ev_ref) > 0, 'the scenario must exercise switches and eliminations'
    eng = RolloutEngine.from_state_dicts(sd_r, sd_c, torch.device('cuda:0'))
    drv = RolloutDriver(eng, x, ei, ea, mask, span=span, global_pos={t: v[:, :2] for t, v in x.items()} if rows == 'morton' else None,
                        edge_threshold=edge_thr, area_threshold=area_thr)
    for s in range(steps):
        drv.step()
        n_pairs, xs, eis = log[s]
        for e in ET:
            assert torch.equal(drv.edge_index[e], eis[e]), (s, e)                    # topology decisions: position for position
        for t in ('joint', 'grain'):
            got = drv._to_caller(t, eng.x[t]).cpu()
            assert rel_err(got, xs[t]) < 1e-4, (s, t, rel_err(got, xs[t]))
    assert drv.grain_event_list == ev_ref
    assert torch.equal(drv.mask['grain'], mask_ref['grain']) and torch.equal(drv.mask['joint'], mask_ref['joint'])
    q = drv.qoi()
    assert q['predicted_grain_events'] == len(ev_ref) and q['switches'] == sum(n for n, _, _ in log)
    # the truth-dependent QoIs: event accounting as test.py:480-491 counts it, layer error through a raster
    truth = {'grain_events': [set()] + [{g + 1 for g in ev_ref[:2]}] * 200, 'imagesize': 501,
             'alpha_pde': lambda frame: ro.plot_polygons(drv.polygons(), 501)}
    drv.truth = truth                                                                # raster: the device kernel (row f4) against PIL's field
    drv.step()
    q = drv.qoi()
    assert q['grain_events_hit_rate'].endswith('/2') and q['last_layer_error'] == 0.0


@pytest.mark.gpu
@pytest.mark.parametrize('name,lxd', [('c1', 40), ('c2', 120)])
def test_area_bookkeeping_matches_the_oracle(name, lxd):
    """graph_trajectory.py:1041-1051, :1100-1103 on the device (row f2, the QoI part) — float64, to rounding of the sums."""
    from graingraphnn_b200.geometry import RegionIndex, area_bookkeeping
    from graingraphnn_b200.graph import build_csr
    x, ei, _ = load_graph(name)
    d = torch.device('cuda:0')
    ng, nj = x['grain'].shape[0], x['joint'].shape[0]
    mask = torch.ones(ng, 1, dtype=torch.int64)                     # the loader's mask is integer (data_loader.py:37-40): the reference's
    mask[::17] = 0                                                  # products are then float64; a few eliminated grains
    x['grain'][:, 4] = torch.rand(ng)
    gj_host = ei[ET[0]][:, mask[ei[ET[0]][0], 0] > 0]                # an eliminated grain has lost its edges (models.py:864-896)
    gj = gj_host.to(d)
    counts, extra, varea = area_bookkeeping(x['grain'].to(d), mask.to(d), build_csr(gj, ng, nj), RegionIndex(gj, ng, nj), lxd)
    ac, ex, va = orc.area_bookkeeping(x['grain'], mask, gj_host, lxd)
    ref = np.full(ng, np.nan)
    for g, v in ac.items():
        ref[g - 1] = v
    # `area * s**2 / area_sum` on a float32 scalar: float64 under the reference's numpy 1.x, float32 under NEP 50 (numpy 2, this
    # image, the oracle run here); the device computes in float64
    np.testing.assert_allclose(counts.cpu().numpy(), ref, rtol=1e-6, equal_nan=True)
    np.testing.assert_allclose(extra.cpu().numpy(), ex, rtol=1e-12)
    refv = np.zeros(nj)
    for j, v in va.items():
        refv[j] = v
    np.testing.assert_allclose(varea.cpu().numpy(), refv, rtol=1e-6)
