#!/usr/bin/env python
"""Row f1 on the device against the host path, on the bench domain (generate-mode, lxd 1320: 1.26e5 grains): crafted predictions
with ~300 switching events and a few eliminations (tests/test_topology_golden._craft), the same inputs through
  host:   D2H of joint rows + predictions, topology.topology_update (position lists), H2D of the new edge lists, set_topology
  device: gg_topology_lists + gg_topology_update + stable compaction + set_topology (DeviceTopology.update)
    python scripts/topology_device_bench.py [--lxd 1320] [--out profiles/r2_topology_device_bench.json]"""
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
import bench  # noqa: E402
from graingraphnn_b200 import topology  # noqa: E402
from graingraphnn_b200.engine import ET_GJ, ET_JG, ET_JJ, RolloutEngine  # noqa: E402
from graingraphnn_b200.topology_device import DeviceTopology  # noqa: E402
from test_topology_golden import _craft  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--lxd', type=int, default=1320)
    ap.add_argument('--switches', type=int, default=300)
    ap.add_argument('--vanish', type=int, default=20)
    ap.add_argument('--cases', type=int, default=3)
    ap.add_argument('--out', default=os.path.join(ROOT, 'profiles', 'r2_topology_device_bench.json'))
    a = ap.parse_args()
    dev = torch.device('cuda:0')
    x, ei, ea, glob, desc = bench.make_domain(a.lxd)
    sd_r, sd_c, _ = bench.synth_weights()
    ng, nj = x['grain'].shape[0], x['joint'].shape[0]
    res = {'domain': desc, 'cases': []}
    for seed in range(4 * a.cases):
        y = _craft(np.random.default_rng(9000 + seed), x, ei, a.switches, a.vanish, 6)
        mask = {'grain': torch.ones(ng, 1), 'joint': torch.ones(nj, 1)}
        # ---- host path on CPU copies (what RolloutDriver(topology='host') does on a step with candidates)
        xh = {k: v.clone() for k, v in x.items()}
        yh = {k: v.clone() for k, v in y.items()}
        prob = torch.sigmoid(yh['edge_event'])
        L1 = ((prob > 0.6) & (ei[ET_JJ][0] < ei[ET_JJ][1])).nonzero().view(-1)
        ge = ((yh['grain_area'] < 1e-4)).nonzero().view(-1)
        ge = ge[torch.argsort(yh['grain_area'][ge])]
        yh['grain_event'] = ge
        act_g, act_j = (yh['grain'][:, 0] > -10).nonzero().view(-1), (yh['joint'][:, 0] > -10).nonzero().view(-1)
        mh = {k: v.clone() for k, v in mask.items()}
        t0 = time.perf_counter()
        try:
            _, ei_host, pairs_host = topology.topology_update(xh, ei, yh, mh, act_g, act_j, threshold=0.6, L1=L1)
        except (KeyError, AssertionError, ValueError, RuntimeError, IndexError):
            continue
        t_host = time.perf_counter() - t0
        # ---- device path: engine resident, candidates in the selection buffers
        eng = RolloutEngine.from_state_dicts(sd_r, sd_c, device=dev)
        eng.set_graph({k: v.to(dev) for k, v in x.items()}, {k: v.to(dev) for k, v in ei.items()}, {k: v.to(dev) for k, v in ea.items()}, global_pos=glob)
        eng.enable_event_selection(mask['grain'])
        sel = eng._events
        pred = {'joint': y['joint'].to(dev).contiguous(), 'grain': y['grain'].to(dev).contiguous()}
        sel.select_edge_events(y['edge_event'].to(dev), ei[ET_JJ].to(dev))
        area = y['grain_area'].to(dev)
        sel.select_grain_events(area.index_select(0, eng.node_order['grain']), eng._event_mask)      # engine rows, as the step selects them
        dt = DeviceTopology(eng, ei, mask)
        dt.profile = True
        dt.edge_prob_fn = lambda v: torch.sigmoid(v.cpu()).to(v.device)          # the host run's own sigmoid: equal probabilities (ties) are the same in both
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        out = dt.update(pred)
        e1.record()
        torch.cuda.synchronize()
        t_dev_wall = time.perf_counter() - t0
        new = dt.edge_index()
        same = all(torch.equal(new[e].cpu(), ei_host[e]) for e in (ET_JJ, ET_JG, ET_GJ)) and torch.equal(out['switching_list'].cpu(), pairs_host)
        same = same and torch.equal(eng.x['joint'].index_select(0, eng._node_rank['joint']).cpu(), xh['joint'])
        res['cases'].append({'grains': ng, 'jj_edges': int(ei[ET_JJ].shape[1]), 'switches': int(pairs_host.shape[0]), 'eliminations': int(out['grain_event'].numel()),
                             'identical_outputs': bool(same), 'host_update_s': t_host, 'device_update_ms_events': e0.elapsed_time(e1),
                             'device_update_wall_s': t_dev_wall, 'device_phases_ms': dt.last_ms,
                             'note': 'device time = lists + sequential kernel + compaction + set_topology (CSR, tile index rebuilt) on the stream; '
                                     'host time = the position-list update alone, without the D2H / H2D of its inputs and outputs'})
        if len(res['cases']) >= a.cases:
            break
    with open(a.out, 'w') as f:
        json.dump(res, f, indent=1)
    print(json.dumps(res)[:3000])


if __name__ == '__main__':
    main()
