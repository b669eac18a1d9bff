#!/usr/bin/env python
"""BASELINE configs 1 and 2 (the reference's own 40x40 and 120x120 graphs: 118 / 1,043 grains) as timed NN steps: the resident
engine replaying its step graph on one B200 next to the reference-order CPU restatement on the host cores, same graph, same seeded
stand-in weights (the shipped .pt files are absent).  These sizes are parity-test cases (tests/), not bench lines: a step is
launch-latency bound on the GPU (~40 kernels of a few microseconds).
    python scripts/configs_bench.py [--out profiles/r2_configs_c1_c2.json]"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tests')):
    sys.path.insert(0, p)
import grain_oracle as orc  # noqa: E402
from util import load_graph  # noqa: E402
from graingraphnn_b200.engine import RolloutEngine  # noqa: E402
from graingraphnn_b200.weights import synth_state_dict  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--out', default=os.path.join(ROOT, 'profiles', 'r2_configs_c1_c2.json'))
    a = ap.parse_args()
    dev = torch.device('cuda:0')
    sd_r, sd_c = synth_state_dict('regressor', 1, head_gain=0.02), synth_state_dict('classifier', 2, head_gain=0.02)
    res = []
    for name, what in (('c1', 'graphs/40_40/seed10020 (config 1)'), ('c2', 'graphs/120_120/seed0 (config 2)')):
        x, ei, ea = load_graph(name)
        edges = sum(int(v.shape[1]) for v in ei.values())
        eng = RolloutEngine.from_state_dicts(sd_r, sd_c, device=dev)
        eng.set_graph({k: v.to(dev) for k, v in x.items()}, {k: v.to(dev) for k, v in ei.items()}, {k: v.to(dev) for k, v in ea.items()})
        for _ in range(3):
            eng.step(6)
        eng.capture(6, warmup=1)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.steps):
            eng.step(6)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.steps
        torch.set_num_threads(os.cpu_count() or 1)
        xc = {k: v.clone() for k, v in x.items()}
        eac = ea
        with torch.no_grad():
            eac = orc.nn_step(sd_r, sd_c, xc, ei, eac, 6)[1]
            t0 = time.perf_counter()
            for _ in range(5):
                eac = orc.nn_step(sd_r, sd_c, xc, ei, eac, 6)[1]
            cpu_ms = (time.perf_counter() - t0) / 5 * 1e3
        res.append({'graph': what, 'grains': int(x['grain'].shape[0]), 'joints': int(x['joint'].shape[0]), 'directed_edges': edges,
                    'gpu_ms_per_step': ms, 'gpu_steps_per_sec': 1e3 / ms, 'gpu_edges_per_sec': edges / ms * 1e3, 'launches_per_step': eng.launches_per_step,
                    'cpu_ms_per_step': cpu_ms, 'cpu_steps_per_sec': 1e3 / cpu_ms, 'cpu_cores': torch.get_num_threads(),
                    'cpu_kind': 'port (oracle/grain_oracle.py, reference op order)'})
        print(res[-1])
    with open(a.out, 'w') as f:
        json.dump({'cases': res, 'weights': 'seeded stand-ins', 'step': 'nn-step, fixed topology, CUDA-graph replay'}, f, indent=1)


if __name__ == '__main__':
    main()
