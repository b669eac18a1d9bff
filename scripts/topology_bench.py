#!/usr/bin/env python
"""Host topology update: the position-list update of graingraphnn_b200.topology beside the O(E)-scan
restatement of the reference's algorithm (oracle/topology_oracle.py — test infrastructure, timed here as the CPU baseline of
this row), on the C2 fixture and on the synthetic bench domain.  CPU only.
    python scripts/topology_bench.py [--patches 36x30] [--out profiles/r1_topology_bench.json]"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tests')):
    sys.path.insert(0, p)
import grain_oracle as orc  # noqa: E402
import topology_oracle as topo  # noqa: E402
from graingraphnn_b200 import topology  # noqa: E402
from test_topology_golden import _craft  # noqa: E402
from util import ET, load_graph  # noqa: E402


def run(fn, x0, ei, y0):
    x = {k: v.clone() for k, v in x0.items()}
    y = {k: v.clone() for k, v in y0.items()}
    orc.regressor_update(x, y, span=0)
    _, y['grain_event'] = orc.event_candidates(y, ei[ET[2]])
    mask = {'grain': torch.ones(x['grain'].shape[0], 1), 'joint': torch.ones(x['joint'].shape[0], 1)}
    active = ((y['grain'][:, 0] > -10).nonzero().view(-1), (y['joint'][:, 0] > -10).nonzero().view(-1))
    t0 = time.perf_counter()
    _, eio, pairs = fn(x, ei, y, mask, *active)
    return time.perf_counter() - t0, eio, pairs, len(y['grain_event'])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--patches', default='36x30')
    ap.add_argument('--out', default=os.path.join(ROOT, 'profiles', 'r1_topology_bench.json'))
    args = ap.parse_args()
    torch.set_num_threads(os.cpu_count())
    out = {'cores': os.cpu_count(), 'cases': []}
    import bench
    cases = [('c2 fixture', load_graph('c2')[:2], 150, 25, 8)]
    from graingraphnn_b200.synth import lattice_domain
    xb, eib, _, pp = lattice_domain(tuple(int(v) for v in args.patches.split('x')))
    cases.append((f'synthetic {pp[0]}x{pp[1]} patches', (xb, eib), 300, 0, 6))
    for label, (x, ei), n_switch, n_vanish, sides in cases:
        for seed in range(20):
            y = _craft(np.random.default_rng(7000 + seed), x, ei, n_switch, n_vanish, sides)
            try:
                t_idx, e1, p1, n_ev = run(topology.topology_update, x, ei, y)
                t_scan, e2, p2, _ = run(topo.topology_update, x, ei, y)
            except (KeyError, AssertionError, ValueError, RuntimeError, IndexError):
                continue
            same = all(torch.equal(e1[k], e2[k]) for k in ET) and torch.equal(p1, p2)
            out['cases'].append({'graph': label, 'grains': x['grain'].shape[0], 'jj_edges': ei[ET[2]].shape[1],
                                 'switches': int(p1.shape[0]), 'eliminations': n_ev, 'identical_outputs': bool(same),
                                 'indexed_update_s': t_idx, 'scan_restatement_s': t_scan, 'ratio': t_scan / t_idx})
            break
    with open(args.out, 'w') as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out))


if __name__ == '__main__':
    main()
