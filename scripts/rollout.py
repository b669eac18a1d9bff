#!/usr/bin/env python
"""A whole rollout (test.py:353-577) on the resident engine: generate (or load) a domain, run the frames, print the QoIs.

    python scripts/rollout.py --lxd 120 --seed 0 [--regressor model/regressor0.pt --classifier model/classifier1.pt]
                              [--truth traj.npz] [--frames 121] [--edge-bias -0.6]

Without the shipped weights it runs on seeded stand-ins (the QoIs are then meaningless, the pipeline is the same); `--truth`
takes an npz with `grain_events` (object array of per-frame id lists, 1-based like traj.grain_events) and `alpha_pde`
([frames, s, s] int), from which the README QoIs (`grain events hit rate`, last-layer error; README.md:64-69) are computed."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from graingraphnn_b200 import generate as G  # noqa: E402
from graingraphnn_b200.engine import RolloutEngine  # noqa: E402
from graingraphnn_b200.rollout import RolloutDriver  # noqa: E402
from graingraphnn_b200.weights import load_weights  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--lxd', type=int, default=40)
    ap.add_argument('--seed', type=int, default=1)
    ap.add_argument('--G', type=float, default=10.0)
    ap.add_argument('--R', type=float, default=2.0)
    ap.add_argument('--span', type=int, default=6)
    ap.add_argument('--frames', type=int, default=121)
    ap.add_argument('--regressor', default=None)
    ap.add_argument('--classifier', default=None)
    ap.add_argument('--truth', default=None)
    ap.add_argument('--head-gain', type=float, default=0.02, help='stand-in weights only: scale of the regressor heads')
    ap.add_argument('--edge-bias', type=float, default=-0.63, help='stand-in weights only: shift of the edge-event logit (untrained logits sit at ~1.0)')
    ap.add_argument('--device', default='cuda:0')
    ap.add_argument('--time-steps', action='store_true', help='synchronise around every step and report the per-step wall times')
    ap.add_argument('--topology', default='device', choices=['device', 'host'], help="where the topology update runs (models.py:614-845)")
    a = ap.parse_args()
    hg = G.generate_graph(lxd=a.lxd, seed=a.seed, G=a.G, R=a.R, span=a.span)
    x, ei, ea, geom = G.model_inputs(hg, a.lxd)
    sd_r, sd_c, wdesc = load_weights(a.regressor, a.classifier, head_gain=a.head_gain)
    if not a.classifier:
        sd_c['lin2.bias'] = sd_c['lin2.bias'] + a.edge_bias
    mask = {k: torch.from_numpy(v) for k, v in hg['mask'].items()}
    truth = None
    if a.truth:
        z = np.load(a.truth, allow_pickle=True)
        truth = {'grain_events': [set(v) for v in z['grain_events']], 'imagesize': int(z['alpha_pde'].shape[1]),
                 'alpha_pde': lambda f: z['alpha_pde'][f], 'train_test_frame_ratio': int(z['ratio']) if 'ratio' in z else 1}
    eng = RolloutEngine.from_state_dicts(sd_r, sd_c, device=a.device)
    drv = RolloutDriver(eng, x, ei, ea, mask, span=a.span, geometry=geom, global_pos=geom['global'], truth=truth, frames=a.frames, lxd=a.lxd, topology=a.topology)
    drv.time_steps = a.time_steps
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    q = drv.run()
    torch.cuda.synchronize()
    q.update({'seconds': time.perf_counter() - t0, 'grains': int(x['grain'].shape[0]), 'weights': wdesc, 'topology': a.topology})
    if a.time_steps:
        q['step_ms'] = [round(t * 1e3, 2) for t in drv.step_seconds]
    print(json.dumps(q))


if __name__ == '__main__':
    main()
