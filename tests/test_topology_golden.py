"""CPU: golden vectors of the reference's OWN topology update — GrainNN_regressor.update, the grain-event
selection of test.py:414-416 and GrainNN_classifier.update run unmodified on crafted predictions (oracle/make_golden_topology.py).
They pin the row that comes next (the topology surgery on the device).  Checked here: the parts that exist — the event
candidates (oracle restatement; the kernel is checked against the same restatement in tests/test_events.py) — the
invariants any implementation of the surgery must keep, and that the geometry feedback (row f2) digests the new topology."""
import os

import numpy as np
import pytest
import torch

import grain_oracle as orc
from util import ET, GOLDEN, load_graph

CASES = [('c1', i) for i in range(5)] + [('c2', i) for i in range(6)]      # gold['c1_cases'], gold['c2_cases']


@pytest.fixture(scope='module')
def gold():
    g = np.load(os.path.join(GOLDEN, 'topology_golden.npz'))
    assert int(g['c1_cases']) == 5 and int(g['c2_cases']) == 6
    return g


def case(gold, name, i):
    k = f'{name}_{i}_'
    return {f[len(k):]: gold[f] for f in gold.files if f.startswith(k)}


@pytest.mark.parametrize('name,i', CASES)
def test_event_candidates_equal_what_the_reference_update_consumed(gold, name, i):
    c = case(gold, name, i)
    _, ei, _ = load_graph(name)
    y = {'edge_event': torch.from_numpy(c['y_edge_event']), 'grain_area': torch.from_numpy(c['y_grain_area'])}
    L1, grain_event = orc.event_candidates(y, ei[ET[2]])
    assert np.array_equal(grain_event.numpy(), c['grain_event_in'])                  # test.py:414-416
    # every switching pair the reference reports (models.py:742) stems from a candidate edge that survived the eliminations
    assert 0 < len(c['switching_list']) <= len(L1)
    assert len(L1) == {('c1', False): 3, ('c1', True): 10, ('c2', False): 12, ('c2', True): 60}[(name, i >= 3)]
    # eliminated grains: the predicted ones, plus any the update forced (models.py:757-759)
    assert np.array_equal(c['grain_event_out'][:len(c['grain_event_in'])], c['grain_event_in'])


@pytest.mark.parametrize('name,i', CASES)
def test_updated_topology_keeps_the_invariants_of_a_periodic_trivalent_tiling(gold, name, i):
    c = case(gold, name, i)
    x, ei, _ = load_graph(name)
    ng, nj = x['grain'].shape[0], x['joint'].shape[0]
    jj, jg, gj = c['ei_jj_out'], c['ei_jg_out'], c['ei_gj_out']
    gone_g = np.unique(c['grain_event_out'])
    live_g = np.setdiff1d(np.arange(ng), gone_g)
    live_j = np.nonzero(c['mask_joint_out'][:, 0] > 0)[0]
    assert np.array_equal(np.nonzero(c['mask_grain_out'][:, 0] == 0)[0], gone_g)
    assert nj - len(live_j) == 2 * len(gone_g) and len(live_j) == 2 * len(live_g)     # each vanished grain takes two joints along
    assert (jj >= 0).all() and (jg >= 0).all()                                         # cleanup dropped the -1 rows (:846-862)
    assert np.array_equal(gj, jg[::-1])                                                # models.py:841
    assert np.array_equal(np.unique(jj[0]), live_j) and np.array_equal(np.unique(jg[0]), live_j)
    assert (np.bincount(jj[0], minlength=nj)[live_j] == 3).all() and (np.bincount(jj[1], minlength=nj)[live_j] == 3).all()
    assert (np.bincount(jg[0], minlength=nj)[live_j] == 3).all()
    assert np.array_equal(np.unique(jg[1]), live_g) and np.bincount(jg[1], minlength=ng)[live_g].min() >= 3
    pairs = set(map(tuple, jj.T.tolist()))
    assert all((b, a) in pairs for a, b in pairs) and len(pairs) == jj.shape[1]        # symmetric, no duplicate edges
    # features: only joints move; the eliminated rows stay in place (masked, not removed)
    assert c['x_joint_out'].shape == (nj, 8) and c['x_grain_out'].shape == (ng, 11)
    assert np.isfinite(c['x_joint_out']).all() and np.isfinite(c['x_grain_out']).all()


@pytest.mark.parametrize('name,i', [('c1', 0), ('c1', 4), ('c2', 0), ('c2', 5)])
def test_geometry_feedback_oracle_on_the_updated_topology(gold, name, i):
    """Row f2 after a topology change: centres of the surviving grains from the surviving joints; vanished grains have none."""
    c = case(gold, name, i)
    ng = c['x_grain_out'].shape[0]
    cen = orc.region_center(torch.from_numpy(c['x_joint_out']), c['ei_gj_out'].astype(np.int64), ng)
    gone = np.unique(c['grain_event_out'])
    assert np.isnan(cen[gone]).all() and not np.isnan(np.delete(cen, gone, axis=0)).any()
    if os.path.exists('/root/reference/graph_trajectory.py') and name == 'c1':          # live reference on this topology
        import sys
        sys.path.insert(0, os.path.join(os.path.dirname(GOLDEN), '..', 'oracle'))
        import make_golden_geometry as mg
        ref = mg.reference_centers(torch.from_numpy(c['x_joint_out'].copy()), torch.from_numpy(c['x_grain_out'].copy()),
                                   c['ei_gj_out'].astype(np.int64), c['ei_jj_out'].astype(np.int64))
        assert np.array_equal(cen, ref, equal_nan=True)


@pytest.mark.parametrize('name,i', CASES)
def test_topology_oracle_equals_the_reference_update(gold, name, i):
    """oracle/topology_oracle.py (restatement of models.py:614-1053) against the reference's own outputs: edge arrays position
    for position, moved joints bit for bit, masks, forced eliminations, the switching list."""
    import topology_oracle as topo
    c = case(gold, name, i)
    x, ei, _ = load_graph(name)
    y = {'joint': torch.from_numpy(c['y_joint'].copy()), 'grain': torch.from_numpy(c['y_grain'].copy()),
         'edge_event': torch.from_numpy(c['y_edge_event']), 'grain_area': torch.from_numpy(c['y_grain_area'])}
    orc.regressor_update(x, y, span=0)                   # models.py:503-516 (test.py:400); the z step (test.py:405-407) follows the update
    _, y['grain_event'] = orc.event_candidates(y, ei[ET[2]])
    mask = {'grain': torch.ones(x['grain'].shape[0], 1), 'joint': torch.ones(x['joint'].shape[0], 1)}
    xo, eio, pairs = topo.topology_update(x, ei, y, mask, torch.from_numpy(c['active_grains']), torch.from_numpy(c['active_joints']))
    for et, short in (((ET[2]), 'jj'), (ET[1], 'jg'), (ET[0], 'gj')):
        assert np.array_equal(eio[et].numpy(), c[f'ei_{short}_out']), short
    assert np.array_equal(pairs.numpy(), c['switching_list'])
    assert np.array_equal(y['grain_event'].numpy(), c['grain_event_out'])
    for t in ('joint', 'grain'):
        assert np.array_equal(xo[t].numpy(), c[f'x_{t}_out']), t
        assert np.array_equal(mask[t].numpy(), c[f'mask_{t}_out']), t
        assert np.array_equal(y[t].numpy(), c[f'y_{t}_out']), t


@pytest.mark.skipif(not os.path.exists('/root/reference/models.py'), reason='reference tree not present')
@pytest.mark.parametrize('name,n_switch,n_vanish,max_sides', [('c1', 25, 5, 7), ('c2', 150, 25, 8), ('c2', 300, 5, 5)])
def test_topology_oracle_equals_the_live_reference_on_denser_event_sets(name, n_switch, n_vanish, max_sides):
    """Only in the build container: the reference's update, imported live on the PyG stub, on event sets denser than the
    fixtures hold (adjacent switching edges, grains of up to 8 sides); where the reference raises on an inconsistent event,
    the oracle must raise the same exception type."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(GOLDEN), '..', 'oracle'))
    import make_golden_topology as mt
    import topology_oracle as topo
    g, x, ei, ea = mt.mgold.load_graph({'c1': '/root/reference/graphs/40_40/seed10020_G1.904_R0.558_span6.pkl',
                                        'c2': '/root/reference/graphs/120_120/seed0_G10.0_R2.0_span6.pkl'}[name], {'c1': 1, 'c2': 3}[name])
    R, C = mt.mgold.build_models(g)
    R.threshold, C.threshold = 1e-4, 0.6
    compared = raised = 0
    for seed in range(4):
        y = mt.craft(np.random.default_rng(9000 + seed), x, ei, n_switch, n_vanish, max_sides)
        ref = ref_exc = None
        try:
            ref = mt.run_reference(R, C, x, ei, ea, y)
        except (KeyError, AssertionError, ValueError, RuntimeError, IndexError) as exc:
            ref_exc = type(exc)
        xo = {k: v.clone() for k, v in x.items()}
        yo = {k: v.clone() for k, v in y.items()}
        orc.regressor_update(xo, yo, span=0)
        _, yo['grain_event'] = orc.event_candidates(yo, ei[ET[2]])
        mask = {'grain': torch.ones(xo['grain'].shape[0], 1), 'joint': torch.ones(xo['joint'].shape[0], 1)}
        active = ((yo['grain'][:, 0] > -10).nonzero().view(-1), (yo['joint'][:, 0] > -10).nonzero().view(-1))   # models.py:505-506
        if ref_exc is not None:
            with pytest.raises(ref_exc):
                topo.topology_update(xo, ei, yo, mask, *active)
            raised += 1
            continue
        _, eio, pairs = topo.topology_update(xo, ei, yo, mask, *active)
        rx, rei, rmask, ry, rpairs, _, _ = ref
        for et in ET:
            assert torch.equal(eio[et], rei[et]), et
        assert torch.equal(pairs, rpairs) and torch.equal(yo['grain_event'], ry['grain_event'])
        for t in ('joint', 'grain'):
            assert torch.equal(xo[t], rx[t]) and torch.equal(mask[t], rmask[t]) and torch.equal(yo[t], ry[t]), t
        compared += 1
    assert compared + raised == 4 and compared >= 1


def _run(fn, name, c):
    x, ei, _ = load_graph(name)
    y = {'joint': torch.from_numpy(c['y_joint'].copy()), 'grain': torch.from_numpy(c['y_grain'].copy()),
         'edge_event': torch.from_numpy(c['y_edge_event']), 'grain_area': torch.from_numpy(c['y_grain_area'])}
    orc.regressor_update(x, y, span=0)
    _, y['grain_event'] = orc.event_candidates(y, ei[ET[2]])
    mask = {'grain': torch.ones(x['grain'].shape[0], 1), 'joint': torch.ones(x['joint'].shape[0], 1)}
    xo, eio, pairs = fn(x, ei, y, mask, torch.from_numpy(c['active_grains']), torch.from_numpy(c['active_joints']))
    return xo, eio, pairs, y, mask


@pytest.mark.parametrize('name,i', CASES)
def test_indexed_host_update_equals_the_reference_update(gold, name, i):
    """graingraphnn_b200.topology (position lists instead of O(E) scans) against the reference's own outputs."""
    from graingraphnn_b200 import topology
    c = case(gold, name, i)
    xo, eio, pairs, y, mask = _run(topology.topology_update, name, c)
    for et, short in ((ET[2], 'jj'), (ET[1], 'jg'), (ET[0], 'gj')):
        assert np.array_equal(eio[et].numpy(), c[f'ei_{short}_out']), short
    assert np.array_equal(pairs.numpy(), c['switching_list'])
    assert np.array_equal(y['grain_event'].numpy(), c['grain_event_out'])
    for t in ('joint', 'grain'):
        assert np.array_equal(xo[t].numpy(), c[f'x_{t}_out']), t
        assert np.array_equal(mask[t].numpy(), c[f'mask_{t}_out']), t
        assert np.array_equal(y[t].numpy(), c[f'y_{t}_out']), t


def _craft(rng, x, ei, n_switch, n_vanish, max_sides):
    """Crafted predictions like oracle/make_golden_topology.craft (kept here so the test needs no reference tree)."""
    nj, ng, E = x['joint'].shape[0], x['grain'].shape[0], ei[ET[2]].shape[1]
    y = {'joint': torch.from_numpy((rng.standard_normal((nj, 2)) * 0.02).astype(np.float32)),
         'grain': torch.from_numpy(np.stack([rng.standard_normal(ng) * 0.02, np.abs(rng.standard_normal(ng)) * 0.01], 1).astype(np.float32))}
    logits = torch.full((E,), -4.0) + torch.from_numpy(rng.standard_normal(E).astype(np.float32)) * 0.3
    fwd = torch.nonzero(ei[ET[2]][0] < ei[ET[2]][1]).view(-1).numpy()
    pick = rng.choice(fwd, n_switch, replace=False)
    logits[torch.from_numpy(pick)] = torch.from_numpy((2.0 + rng.random(n_switch) * 2).astype(np.float32))
    y['edge_event'] = logits
    area = x['grain'][:, 3] + torch.tanh(y['grain'][:, 0]) / 20
    deg = torch.bincount(ei[ET[1]][1], minlength=ng)
    small = torch.nonzero(deg <= max_sides).view(-1).numpy()
    vanish = rng.choice(small, min(n_vanish, len(small)), replace=False)
    area[torch.from_numpy(vanish)] = torch.from_numpy((rng.random(len(vanish)) * 9e-5).astype(np.float32))
    y['grain_area'] = area
    return y


@pytest.mark.parametrize('name,n_switch,n_vanish,max_sides', [('c1', 25, 5, 7), ('c1', 60, 12, 8), ('c2', 150, 25, 8), ('c2', 400, 60, 8)])
def test_indexed_host_update_equals_the_oracle_on_denser_event_sets(name, n_switch, n_vanish, max_sides):
    """Adjacent switching edges, grains of up to 8 sides, forced eliminations: the position-list update and the O(E)-scan
    oracle (itself pinned by the reference's outputs) agree array for array, and raise the same exception type on event sets
    the reference's algorithm cannot digest."""
    import topology_oracle as topo
    from graingraphnn_b200 import topology
    x0, ei, _ = load_graph(name)
    agreed = raised = 0
    for seed in range(5):
        y0 = _craft(np.random.default_rng(7000 + seed), x0, ei, n_switch, n_vanish, max_sides)
        res = []
        for fn in (topo.topology_update, topology.topology_update):
            x = {k: v.clone() for k, v in x0.items()}
            y = {k: v.clone() for k, v in y0.items()}
            orc.regressor_update(x, y, span=0)
            _, y['grain_event'] = orc.event_candidates(y, ei[ET[2]])
            mask = {'grain': torch.ones(x['grain'].shape[0], 1), 'joint': torch.ones(x['joint'].shape[0], 1)}
            active = ((y['grain'][:, 0] > -10).nonzero().view(-1), (y['joint'][:, 0] > -10).nonzero().view(-1))
            try:
                _, eio, pairs = fn(x, ei, y, mask, *active)
                res.append((x, eio, pairs, y, mask))
            except (KeyError, AssertionError, ValueError, RuntimeError, IndexError) as exc:
                res.append(type(exc))
        a, b = res
        if isinstance(a, type) or isinstance(b, type):
            assert a is b, (a, b)
            raised += 1
            continue
        for et in ET:
            assert torch.equal(a[1][et], b[1][et]), et
        assert torch.equal(a[2], b[2]) and torch.equal(a[3]['grain_event'], b[3]['grain_event'])
        for t in ('joint', 'grain'):
            assert torch.equal(a[0][t], b[0][t]) and torch.equal(a[4][t], b[4][t]) and torch.equal(a[3][t], b[3][t]), t
        agreed += 1
    assert agreed >= 1 and agreed + raised == 5


@pytest.mark.parametrize('name,i', [('c1', 3), ('c2', 0), ('c2', 4)])
def test_classifier_update_has_the_reference_signature_and_results(gold, name, i):
    """models.GrainNN_classifier.update(x_dict, edge_index_dict, edge_attr, y_dict, mask, geometry_scaling, nucleation_prob)
    — the call of test.py:426 — on host tensors: the reference's outputs, in-place updates included."""
    from graingraphnn_b200.engine import _Hyper
    from graingraphnn_b200.models import GrainNN_classifier
    c = case(gold, name, i)
    x, ei, ea = load_graph(name)
    C = GrainNN_classifier(_Hyper({'grain': list(range(11)), 'joint': list(range(8))}, {'grain': [0, 1], 'joint': [0, 1]}, 96,
                                  (['grain', 'joint', 'mask'], list(ET)), 'cpu'))
    C.threshold = 0.6                                                          # test.py:187
    y = {'joint': torch.from_numpy(c['y_joint'].copy()), 'grain': torch.from_numpy(c['y_grain'].copy()),
         'edge_event': torch.from_numpy(c['y_edge_event']), 'grain_area': torch.from_numpy(c['y_grain_area'])}
    orc.regressor_update(x, y, span=0)
    _, y['grain_event'] = orc.event_candidates(y, ei[ET[2]])
    mask = {'grain': torch.ones(x['grain'].shape[0], 1), 'joint': torch.ones(x['joint'].shape[0], 1)}
    gs = {'domain_offset': 0, 'domain_factor': 1, 'active_grains': torch.from_numpy(c['active_grains']),
          'active_joints': torch.from_numpy(c['active_joints'])}
    xo, eio, pairs = C.update(x, ei, ea, y, mask, gs, 0.0)
    assert xo is x and eio is ei
    for et, short in ((ET[2], 'jj'), (ET[1], 'jg'), (ET[0], 'gj')):
        assert np.array_equal(ei[et].numpy(), c[f'ei_{short}_out']), short
    assert np.array_equal(pairs.numpy(), c['switching_list']) and np.array_equal(y['grain_event'].numpy(), c['grain_event_out'])
    for t in ('joint', 'grain'):
        assert np.array_equal(x[t].numpy(), c[f'x_{t}_out']) and np.array_equal(mask[t].numpy(), c[f'mask_{t}_out'])
        assert np.array_equal(y[t].numpy(), c[f'y_{t}_out'])


def _craft_on(rng, x, ei, mask, n_switch, n_vanish, max_sides):
    """Crafted events on an EVOLVING topology: only live grains vanish, every other area stays well above the threshold."""
    nj, ng, E = x['joint'].shape[0], x['grain'].shape[0], ei[ET[2]].shape[1]
    y = {'joint': torch.from_numpy((rng.standard_normal((nj, 2)) * 0.01).astype(np.float32)),
         'grain': torch.from_numpy(np.stack([rng.standard_normal(ng) * 0.02, np.abs(rng.standard_normal(ng)) * 0.01], 1).astype(np.float32))}
    logits = torch.full((E,), -4.0)
    fwd = torch.nonzero(ei[ET[2]][0] < ei[ET[2]][1]).view(-1).numpy()
    pick = rng.choice(fwd, min(n_switch, len(fwd)), replace=False)
    logits[torch.from_numpy(pick)] = torch.from_numpy((2.0 + rng.random(len(pick)) * 2).astype(np.float32))
    y['edge_event'] = logits
    area = (x['grain'][:, 3] + torch.tanh(y['grain'][:, 0]) / 20).clamp(min=1e-3)
    deg = torch.bincount(ei[ET[1]][1], minlength=ng)
    small = torch.nonzero((deg <= max_sides) & (deg > 0) & (mask['grain'][:, 0] > 0)).view(-1).numpy()
    van = rng.choice(small, min(n_vanish, len(small)), replace=False)
    area[torch.from_numpy(van)] = torch.from_numpy((rng.random(len(van)) * 9e-5).astype(np.float32))
    y['grain_area'] = area
    return y


@pytest.mark.skipif(not os.path.exists('/root/reference/models.py'), reason='reference tree not present')
@pytest.mark.parametrize('name,chain,n_switch,n_vanish,forced_expected', [('c1', 0, 8, 3, 0), ('c2', 0, 80, 25, 2), ('c2', 3, 80, 25, 1)])
def test_six_consecutive_updates_equal_the_live_reference(name, chain, n_switch, n_vanish, forced_expected):
    """Only in the build container: six updates in a row on an evolving topology (triangles appear, masks carry over) —
    the only vectors on which the reference's forced eliminations (models.py:967-973, :757-759) fire.  Each round the
    product (`topology.py`) and the oracle start from the reference's state and must reproduce its next state exactly."""
    import contextlib
    import io
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(GOLDEN), '..', 'oracle'))
    import make_golden_topology as mt
    import topology_oracle as topo
    from graingraphnn_b200 import topology
    g, x, ei, ea = mt.mgold.load_graph({'c1': '/root/reference/graphs/40_40/seed10020_G1.904_R0.558_span6.pkl',
                                        'c2': '/root/reference/graphs/120_120/seed0_G10.0_R2.0_span6.pkl'}[name], {'c1': 1, 'c2': 3}[name])
    R, C = mt.mgold.build_models(g)
    R.threshold, C.threshold = 1e-4, 0.6
    rng = np.random.default_rng(500 + chain)
    mask = {'grain': torch.ones(x['grain'].shape[0], 1), 'joint': torch.ones(x['joint'].shape[0], 1)}
    forced = 0
    for rnd in range(6):
        y = _craft_on(rng, x, ei, mask, n_switch, n_vanish, 6)
        rx, rei, ry, rm = ({k: v.clone() for k, v in d.items()} for d in (x, ei, y, mask))
        gs = {'domain_offset': 0, 'domain_factor': 1}
        with contextlib.redirect_stdout(io.StringIO()):
            R.update(rx, ry, gs)                                                                   # test.py:400
            ry['grain_event'] = ((rm['grain'][:, 0] > 0) & (ry['grain_area'] < R.threshold)).nonzero().view(-1)
            ry['grain_event'] = ry['grain_event'][torch.argsort(ry['grain_area'][ry['grain_event']])]
            n_in = len(ry['grain_event'])
            rx, rei, rpairs = C.update(rx, rei, ea, ry, rm, gs, 0.0)                               # test.py:426
        for fn in (topology.topology_update, topo.topology_update):
            ox, oy, om = ({k: v.clone() for k, v in d.items()} for d in (x, y, mask))
            orc.regressor_update(ox, oy, span=0)
            _, oy['grain_event'] = orc.event_candidates(oy, ei[ET[2]], om['grain'])
            _, oei, op = fn(ox, ei, oy, om, gs['active_grains'], gs['active_joints'])
            for et in ET:
                assert torch.equal(oei[et], rei[et]), (rnd, et)
            assert torch.equal(op, rpairs) and torch.equal(oy['grain_event'], ry['grain_event']), rnd
            for t in ('joint', 'grain'):
                assert torch.equal(ox[t], rx[t]) and torch.equal(om[t], rm[t]) and torch.equal(oy[t], ry[t]), (rnd, t)
        forced += len(ry['grain_event']) - n_in
        x, ei, mask = rx, rei, rm
    assert forced == forced_expected


@pytest.mark.parametrize('impl', ['product', 'oracle'])
def test_forced_eliminations_equal_the_reference(gold, impl):
    """The committed vector on which the reference's forced eliminations fire (two grains beyond the predicted ones): state
    after three earlier updates on C2, masks carried over."""
    import topology_oracle as topo
    from graingraphnn_b200 import topology
    k = 'c2_forced_'
    c = {f[len(k):]: gold[f] for f in gold.files if f.startswith(k)}
    x = {t: torch.from_numpy(c[f'x_{t}_in'].copy()) for t in ('joint', 'grain')}
    mask = {t: torch.from_numpy(c[f'mask_{t}_in'].copy()) for t in ('joint', 'grain')}
    ei = {et: torch.from_numpy(c[f'ei_{short}_in'].astype(np.int64)) for et, short in ((ET[0], 'gj'), (ET[1], 'jg'), (ET[2], 'jj'))}
    y = {'joint': torch.from_numpy(c['y_joint'].copy()), 'grain': torch.from_numpy(c['y_grain'].copy()),
         'edge_event': torch.from_numpy(c['y_edge_event']), 'grain_area': torch.from_numpy(c['y_grain_area'])}
    orc.regressor_update(x, y, span=0)
    _, y['grain_event'] = orc.event_candidates(y, ei[ET[2]], mask['grain'])
    assert np.array_equal(y['grain_event'].numpy(), c['grain_event_in'])
    fn = topology.topology_update if impl == 'product' else topo.topology_update
    _, eio, pairs = fn(x, ei, y, mask, torch.from_numpy(c['active_grains']), torch.from_numpy(c['active_joints']))
    assert len(c['grain_event_out']) == len(c['grain_event_in']) + 2
    assert np.array_equal(y['grain_event'].numpy(), c['grain_event_out']) and np.array_equal(pairs.numpy(), c['switching_list'])
    for et, short in ((ET[2], 'jj'), (ET[1], 'jg'), (ET[0], 'gj')):
        assert np.array_equal(eio[et].numpy(), c[f'ei_{short}_out']), short
    for t in ('joint', 'grain'):
        assert np.array_equal(x[t].numpy(), c[f'x_{t}_out']) and np.array_equal(mask[t].numpy(), c[f'mask_{t}_out'])
        assert np.array_equal(y[t].numpy(), c[f'y_{t}_out'])


def test_update_from_device_selected_candidates_equals_internal_selection(gold):
    """`L1` handed in (what EventSelector.fetch() returns: ascending edge ids) instead of being selected from the full
    edge_event array inside the update: same result."""
    from graingraphnn_b200 import topology
    c = case(gold, 'c2', 4)
    outs = []
    for hand_in in (False, True):
        x, ei, _ = load_graph('c2')
        y = {'joint': torch.from_numpy(c['y_joint'].copy()), 'grain': torch.from_numpy(c['y_grain'].copy()),
             'edge_event': torch.from_numpy(c['y_edge_event']), 'grain_area': torch.from_numpy(c['y_grain_area'])}
        orc.regressor_update(x, y, span=0)
        L1, y['grain_event'] = orc.event_candidates(y, ei[ET[2]])
        mask = {'grain': torch.ones(x['grain'].shape[0], 1), 'joint': torch.ones(x['joint'].shape[0], 1)}
        _, eio, pairs = topology.topology_update(x, ei, y, mask, torch.from_numpy(c['active_grains']),
                                                 torch.from_numpy(c['active_joints']), L1=L1 if hand_in else None)
        outs.append((eio, pairs, x))
    assert all(torch.equal(outs[0][0][et], outs[1][0][et]) for et in ET) and torch.equal(outs[0][1], outs[1][1])
    assert torch.equal(outs[0][2]['joint'], outs[1][2]['joint'])
    assert np.array_equal(outs[1][0][ET[2]].numpy(), c['ei_jj_out'])


@pytest.mark.skipif(not os.path.exists('/root/reference/models.py'), reason='reference tree not present')
@pytest.mark.parametrize('name,prob,seed', [('c1', 0.03, 1), ('c2', 0.01, 2), ('c2', 0.004, 3)])
def test_nucleation_equals_the_live_reference(name, prob, seed):
    """Only in the build container: the optional nucleation branch (models.py:771-835) with the same torch generator state —
    new grains / junctions, re-bound feature and mask tensors, edge arrays position for position."""
    import contextlib
    import io
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(GOLDEN), '..', 'oracle'))
    import make_golden_topology as mt
    from graingraphnn_b200 import topology
    g, x, ei, ea = mt.mgold.load_graph({'c1': '/root/reference/graphs/40_40/seed10020_G1.904_R0.558_span6.pkl',
                                        'c2': '/root/reference/graphs/120_120/seed0_G10.0_R2.0_span6.pkl'}[name], {'c1': 1, 'c2': 3}[name])
    R, C = mt.mgold.build_models(g)
    R.threshold, C.threshold = 1e-4, 0.6
    mask = {'grain': torch.ones(x['grain'].shape[0], 1), 'joint': torch.ones(x['joint'].shape[0], 1)}
    y = _craft_on(np.random.default_rng(seed), x, ei, mask, 10, 3, 6)
    rx, rei, ry, rm = ({k: v.clone() for k, v in d.items()} for d in (x, ei, y, mask))
    gs = {'domain_offset': 0, 'domain_factor': 1}
    with contextlib.redirect_stdout(io.StringIO()):
        R.update(rx, ry, gs)
        ry['grain_event'] = ((rm['grain'][:, 0] > 0) & (ry['grain_area'] < R.threshold)).nonzero().view(-1)
        ry['grain_event'] = ry['grain_event'][torch.argsort(ry['grain_area'][ry['grain_event']])]
        torch.manual_seed(seed)
        rx, rei, rpairs = C.update(rx, rei, ea, ry, rm, gs, prob)
    ox, oy, om = ({k: v.clone() for k, v in d.items()} for d in (x, y, mask))
    orc.regressor_update(ox, oy, span=0)
    _, oy['grain_event'] = orc.event_candidates(oy, ei[ET[2]], om['grain'])
    torch.manual_seed(seed)
    ox, oei, op = topology.topology_update(ox, ei, oy, om, gs['active_grains'], gs['active_joints'], nucleation_prob=prob)
    assert rx['grain'].shape[0] > x['grain'].shape[0]                       # grains were born
    assert rx['joint'].shape[0] - x['joint'].shape[0] == 2 * (rx['grain'].shape[0] - x['grain'].shape[0])
    for et in ET:
        assert torch.equal(oei[et], rei[et]), et
    assert torch.equal(op, rpairs)
    for t in ('joint', 'grain'):
        assert torch.equal(ox[t], rx[t]) and torch.equal(om[t], rm[t]) and om[t].dtype == rm[t].dtype, t


def test_nucleation_equals_the_reference_vector(gold):
    """Committed vector of the nucleation branch (C1, torch.manual_seed before the update): new grains and junctions, grown
    feature / mask tensors, edge arrays position for position."""
    from graingraphnn_b200 import topology
    k = 'c1_nucl_'
    c = {f[len(k):]: gold[f] for f in gold.files if f.startswith(k)}
    x, ei, _ = load_graph('c1')
    y = {'joint': torch.from_numpy(c['y_joint'].copy()), 'grain': torch.from_numpy(c['y_grain'].copy()),
         'edge_event': torch.from_numpy(c['y_edge_event']), 'grain_area': torch.from_numpy(c['y_grain_area'])}
    orc.regressor_update(x, y, span=0)
    mask = {'grain': torch.ones(x['grain'].shape[0], 1), 'joint': torch.ones(x['joint'].shape[0], 1)}
    _, y['grain_event'] = orc.event_candidates(y, ei[ET[2]], mask['grain'])
    torch.manual_seed(int(c['seed']))
    x, eio, pairs = topology.topology_update(x, ei, y, mask, torch.from_numpy(c['active_grains']), torch.from_numpy(c['active_joints']),
                                             nucleation_prob=float(c['prob']))
    assert x['grain'].shape[0] > 118 and x['joint'].shape[0] - 236 == 2 * (x['grain'].shape[0] - 118)
    for et, short in ((ET[2], 'jj'), (ET[1], 'jg'), (ET[0], 'gj')):
        assert np.array_equal(eio[et].numpy(), c[f'ei_{short}_out']), short
    assert np.array_equal(pairs.numpy(), c['switching_list'])
    for t in ('joint', 'grain'):
        assert np.array_equal(x[t].numpy(), c[f'x_{t}_out']) and np.array_equal(mask[t].numpy(), c[f'mask_{t}_out'])
