"""Generate tests/golden/topology_golden.npz: inputs and outputs of the REFERENCE's own topology update —
GrainNN_regressor.update (models.py:473-516), the grain-event selection of test.py:414-416 and GrainNN_classifier.update
(models.py:614-845: grain elimination, neighbour switching, cleanup) — imported unmodified from /root/reference and run on
oracle/pyg_stub.  Run once in the build container:  python oracle/make_golden_topology.py

TEST INFRASTRUCTURE ONLY.  These vectors pin the NEXT row to be built (the topology surgery on the
device); today they are consumed by tests/test_topology_golden.py, which checks the parts of the row that exist — the event
candidates (gg_select_events) — and the invariants any implementation must keep.  Predictions are crafted (a few vanishing
grains, a few switching edges, small joint motions): the shipped weights are absent and seeded stand-in weights put half
of all edges above the threshold, which the reference's update cannot digest.  Seeds on which the reference itself raises
(KeyError / AssertionError on geometrically inconsistent events) are skipped and counted.
"""
import contextlib
import io
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, 'pyg_stub'))
import ref_shims  # noqa: E402

ref_shims.install_plot_stubs()
ref_shims.add_reference_to_path()
import grain_oracle as orc  # noqa: E402
import make_golden as mgold  # noqa: E402  (load_graph / build_models of the NN golden script)

OUT = os.path.join(HERE, '..', 'tests', 'golden')
ET = mgold.ET
GJ, JG, JJ = ET


def craft(rng, x, ei, n_switch, n_vanish, max_sides=5):
    nj, ng, E = x['joint'].shape[0], x['grain'].shape[0], ei[JJ].shape[1]
    y = {'joint': torch.from_numpy((rng.standard_normal((nj, 2)) * 0.02).astype(np.float32)),
         'grain': torch.from_numpy(np.stack([rng.standard_normal(ng) * 0.02, np.abs(rng.standard_normal(ng)) * 0.01], 1).astype(np.float32))}
    logits = torch.full((E,), -4.0) + torch.from_numpy(rng.standard_normal(E).astype(np.float32)) * 0.3
    fwd = torch.nonzero(ei[JJ][0] < ei[JJ][1]).view(-1).numpy()
    pick = rng.choice(fwd, n_switch, replace=False)
    logits[torch.from_numpy(pick)] = torch.from_numpy((2.0 + rng.random(n_switch) * 2).astype(np.float32))
    y['edge_event'] = logits
    area = x['grain'][:, 3] + torch.tanh(y['grain'][:, 0]) / 20                  # models.py:445
    deg = torch.bincount(ei[JG][1], minlength=ng)
    small = torch.nonzero(deg <= max_sides).view(-1).numpy()
    vanish = rng.choice(small, min(n_vanish, len(small)), replace=False)
    area[torch.from_numpy(vanish)] = torch.from_numpy((rng.random(len(vanish)) * 9e-5).astype(np.float32))
    y['grain_area'] = area
    return y


@torch.no_grad()
def run_reference(R, C, x, ei, ea, y):
    x = {k: v.clone() for k, v in x.items()}
    ei = {k: v.clone() for k, v in ei.items()}
    y = {k: v.clone() for k, v in y.items()}
    mask = {'grain': torch.ones(x['grain'].shape[0], 1), 'joint': torch.ones(x['joint'].shape[0], 1)}
    gs = {'domain_offset': 0, 'domain_factor': 1}
    with contextlib.redirect_stdout(io.StringIO()):
        R.update(x, y, gs)                                                       # test.py:400
        y['grain_event'] = ((mask['grain'][:, 0] > 0) & (y['grain_area'] < R.threshold)).nonzero().view(-1)     # test.py:414
        y['grain_event'] = y['grain_event'][torch.argsort(y['grain_area'][y['grain_event']])]                  # :416
        first_events = y['grain_event'].clone()
        x, ei, pairs = C.update(x, ei, ea, y, mask, gs, 0.0)                     # test.py:424
    return x, ei, mask, y, pairs, first_events, gs


def forced_case(gold):
    """The one committed vector on which the reference's FORCED eliminations fire (models.py:967-973, :757-759): round 3 of a
    chain of updates on the C2 graph (tests/test_topology_golden.py::test_six_consecutive_updates_equal_the_live_reference,
    chain 0) — state before the round (features, edges, masks), the crafted predictions, the reference's state after it."""
    sys.path.insert(0, os.path.join(HERE, '..', 'tests'))
    from test_topology_golden import _craft_on
    g, x, ei, ea = mgold.load_graph('/root/reference/graphs/120_120/seed0_G10.0_R2.0_span6.pkl', 3)
    R, C = mgold.build_models(g)
    R.threshold, C.threshold = 1e-4, 0.6
    rng = np.random.default_rng(500)
    mask = {'grain': torch.ones(x['grain'].shape[0], 1), 'joint': torch.ones(x['joint'].shape[0], 1)}
    for rnd in range(4):
        y = _craft_on(rng, x, ei, mask, 80, 25, 6)
        rx, rei, ry, rm = ({k: v.clone() for k, v in d.items()} for d in (x, ei, y, mask))
        gs = {'domain_offset': 0, 'domain_factor': 1}
        with contextlib.redirect_stdout(io.StringIO()):
            R.update(rx, ry, gs)
            ry['grain_event'] = ((rm['grain'][:, 0] > 0) & (ry['grain_area'] < R.threshold)).nonzero().view(-1)
            ry['grain_event'] = ry['grain_event'][torch.argsort(ry['grain_area'][ry['grain_event']])]
            first = ry['grain_event'].clone()
            rx, rei, pairs = C.update(rx, rei, ea, ry, rm, gs, 0.0)
        if rnd == 3:
            k = 'c2_forced'
            for t in ('joint', 'grain'):
                gold[f'{k}_x_{t}_in'], gold[f'{k}_mask_{t}_in'] = x[t].numpy(), mask[t].numpy()
                gold[f'{k}_y_{t}'], gold[f'{k}_x_{t}_out'] = y[t].numpy(), rx[t].numpy()
                gold[f'{k}_mask_{t}_out'], gold[f'{k}_y_{t}_out'] = rm[t].numpy(), ry[t].numpy()
            gold[f'{k}_y_edge_event'], gold[f'{k}_y_grain_area'] = y['edge_event'].numpy(), y['grain_area'].numpy()
            gold[f'{k}_grain_event_in'], gold[f'{k}_grain_event_out'] = first.numpy(), ry['grain_event'].numpy()
            gold[f'{k}_switching_list'] = pairs.numpy()
            gold[f'{k}_active_grains'], gold[f'{k}_active_joints'] = gs['active_grains'].numpy(), gs['active_joints'].numpy()
            for et, short in mgold.SHORT.items():
                gold[f'{k}_ei_{short}_in'] = ei[et].numpy().astype(np.int32)
                gold[f'{k}_ei_{short}_out'] = rei[et].numpy().astype(np.int32)
            print('forced case: events in/out', len(first), len(ry['grain_event']), 'switches', len(pairs))
            assert len(ry['grain_event']) > len(first)
        x, ei, mask = rx, rei, rm


def nucleation_case(gold, prob=0.03, seed=1):
    """The optional nucleation branch (models.py:771-835) on the C1 graph: torch.manual_seed(seed) right before the update,
    so the draws (one vector over the junctions, two angles per site) are reproducible wherever the test runs."""
    sys.path.insert(0, os.path.join(HERE, '..', 'tests'))
    from test_topology_golden import _craft_on
    g, x, ei, ea = mgold.load_graph('/root/reference/graphs/40_40/seed10020_G1.904_R0.558_span6.pkl', 1)
    R, C = mgold.build_models(g)
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
        rx, rei, pairs = C.update(rx, rei, ea, ry, rm, gs, prob)
    k = 'c1_nucl'
    gold[f'{k}_prob'], gold[f'{k}_seed'] = np.array(prob), np.array(seed)
    for t in ('joint', 'grain'):
        gold[f'{k}_y_{t}'], gold[f'{k}_x_{t}_out'], gold[f'{k}_mask_{t}_out'] = y[t].numpy(), rx[t].numpy(), rm[t].numpy()
    gold[f'{k}_y_edge_event'], gold[f'{k}_y_grain_area'] = y['edge_event'].numpy(), y['grain_area'].numpy()
    gold[f'{k}_switching_list'] = pairs.numpy()
    gold[f'{k}_active_grains'], gold[f'{k}_active_joints'] = gs['active_grains'].numpy(), gs['active_joints'].numpy()
    for et, short in mgold.SHORT.items():
        gold[f'{k}_ei_{short}_out'] = rei[et].numpy().astype(np.int32)
    print('nucleation case: grains', x['grain'].shape[0], '->', rx['grain'].shape[0], 'joints', x['joint'].shape[0], '->', rx['joint'].shape[0])
    assert rx['grain'].shape[0] > x['grain'].shape[0]


def main():
    gold, log = {}, {}
    plans = {'c1': (1, [(3, 1, 5, 3), (10, 3, 6, 2)]),                  # (switching edges, vanishing grains, max sides, cases)
             'c2': (3, [(12, 4, 5, 3), (60, 10, 7, 3)])}
    for name, (factor, configs) in plans.items():
        g, x, ei, ea = mgold.load_graph({'c1': '/root/reference/graphs/40_40/seed10020_G1.904_R0.558_span6.pkl',
                                         'c2': '/root/reference/graphs/120_120/seed0_G10.0_R2.0_span6.pkl'}[name], factor)
        R, C = mgold.build_models(g)
        R.threshold, C.threshold = 1e-4, 0.6                                     # test.py:186-187
        done, tried, seed = 0, 0, 0
        for n_switch, n_vanish, max_sides, want in configs:
          target = done + want
          while done < target and tried < 200:
            rng = np.random.default_rng(5000 + seed)
            seed += 1
            tried += 1
            y = craft(rng, x, ei, n_switch, n_vanish, max_sides)
            try:
                xo, eio, mask, yo, pairs, first_events, gs = run_reference(R, C, x, ei, ea, y)
            except (KeyError, AssertionError, ValueError, RuntimeError, IndexError) as exc:
                log.setdefault(name, []).append(f'seed {seed - 1}: {type(exc).__name__}')
                continue
            k = f'{name}_{done}'
            for t in ('joint', 'grain'):
                gold[f'{k}_y_{t}'] = y[t].numpy()
                gold[f'{k}_x_{t}_out'] = xo[t].numpy()
                gold[f'{k}_mask_{t}_out'] = mask[t].numpy()
                gold[f'{k}_y_{t}_out'] = yo[t].numpy()
            gold[f'{k}_y_edge_event'] = y['edge_event'].numpy()
            gold[f'{k}_y_grain_area'] = y['grain_area'].numpy()
            gold[f'{k}_grain_event_in'] = first_events.numpy()
            gold[f'{k}_grain_event_out'] = yo['grain_event'].numpy()
            gold[f'{k}_switching_list'] = pairs.numpy()
            gold[f'{k}_active_grains'] = gs['active_grains'].numpy()
            gold[f'{k}_active_joints'] = gs['active_joints'].numpy()
            for et, short in mgold.SHORT.items():
                gold[f'{k}_ei_{short}_out'] = eio[et].numpy().astype(np.int32)
            print(k, 'events in/out', len(first_events), len(yo['grain_event']), 'switches', len(pairs))
            done += 1
        gold[f'{name}_cases'] = np.array(done)
        print(name, 'cases', done, 'tried', tried, 'reference raised on', log.get(name, []))
    forced_case(gold)
    nucleation_case(gold)
    np.savez_compressed(os.path.join(OUT, 'topology_golden.npz'), **gold)
    print({k: v.shape for k, v in gold.items() if k.endswith('_0_ei_jj_out') or k.endswith('switching_list') or 'grain_event' in k})


if __name__ == '__main__':
    main()
