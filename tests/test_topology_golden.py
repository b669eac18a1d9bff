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

CASES = [(n, i) for n in ('c1', 'c2') for i in range(3)]


@pytest.fixture(scope='module')
def gold():
    return np.load(os.path.join(GOLDEN, 'topology_golden.npz'))


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
    assert len(L1) == {'c1': 3, 'c2': 12}[name]
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


@pytest.mark.parametrize('name,i', [('c1', 0), ('c2', 0), ('c2', 2)])
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
