#!/usr/bin/env python
"""Geometry feedback at the bench size: device time of gg_region_center (CUDA events), its HBM
roofline, the rollout step with and without it (graph replay), and the reference-order host port beside it.
    python scripts/geometry_bench.py [--patches 36x30] [--out gpurun_out/geometry_bench.json]"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import bench  # noqa: E402
from graingraphnn_b200.engine import ET_GJ, RolloutEngine  # noqa: E402
from graingraphnn_b200.geometry import RegionIndex, region_center  # noqa: E402


def events(fn, reps, flush=None):
    ts = []
    for _ in range(reps):
        if flush is not None:
            flush.add_(1.0)                           # > L2: the next launch reads from HBM
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--patches', default='36x30')
    ap.add_argument('--out', default='gpurun_out/geometry_bench.json')
    args = ap.parse_args()
    patches = tuple(int(v) for v in args.patches.split('x'))
    dev = torch.device('cuda:0')
    from graingraphnn_b200.synth import lattice_domain
    x, ei, glob, pp = lattice_domain(patches or (36, 30))
    ng, nj, E = x['grain'].shape[0], x['joint'].shape[0], ei[ET_GJ].shape[1]
    xj, xg = x['joint'].to(dev), x['grain'].to(dev)
    idx = RegionIndex(ei[ET_GJ].to(dev), ng, nj)
    centers = torch.empty(ng, 2, dtype=torch.float64, device=dev)
    flush = torch.zeros(64 << 20, device=dev)         # 256 MB
    for _ in range(3):
        region_center(xj, idx, xg, centers=centers)
    t_kernel = events(lambda: region_center(xj, idx, xg, centers=centers), 20, flush)
    t_warm = events(lambda: region_center(xj, idx, xg, centers=centers), 20)
    # algorithmic bytes per launch: per edge key + col (8 B) and a joint (x, y) (8 B, every joint row is needed once per
    # incident grain but lives in a 32-B sector with its unused columns: 32 B per joint of DRAM traffic at best);
    # per grain rowptr (4 B), centre out (16 B), (x, y) write-back (8 B)
    t_bykey = events(lambda: region_center(xj, idx, xg, centers=centers, presorted=False), 20, flush)
    alg = E * 4 + nj * 8 + ng * (4 + 16 + 8)            # joint ids in dict order (gg_region_sort, once per topology): no keys
    peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json'))) if os.path.exists(os.path.join(ROOT, 'MEASURED_PEAKS.json')) else {}
    out = {'grains': ng, 'joints': nj, 'edges_gj': E, 'region_center_us_cold_l2': t_kernel * 1e3, 'region_center_us_warm_l2': t_warm * 1e3, 'region_center_by_key_us_cold_l2': t_bykey * 1e3,
           'algorithmic_bytes': alg, 'achieved_GBps_cold': alg / (t_kernel * 1e-3) / 1e9, 'peaks_file': peaks}
    # event candidates (row f1, first stage): one pass over the jj edge logits + one over the grain areas
    from graingraphnn_b200.events import EventSelector
    from graingraphnn_b200.engine import ET_JJ
    jj = ei[ET_JJ].to(dev)
    logits = torch.randn(jj.shape[1], device=dev) * 1.5 - 4.0
    area = torch.rand(ng, device=dev) * 0.02
    sel = EventSelector(dev)

    def select():
        sel.select_edge_events(logits, jj)
        sel.select_grain_events(area)
    for _ in range(3):
        select()
    t_sel = events(select, 20, flush)
    ev = sel.fetch()
    out['select_events_us_cold_l2'] = t_sel * 1e3
    out['select_events_algorithmic_bytes'] = 4 * jj.shape[1] + 4 * ng
    out['select_events_GBps'] = out['select_events_algorithmic_bytes'] / (t_sel * 1e-3) / 1e9
    out['select_events_candidates'] = [len(ev['L1']), len(ev['grain_event'])]
    out['select_events_d2h_bytes_vs_full'] = [sel.d2h_bytes, 4 * (jj.shape[1] + 3 * ng + 2 * nj)]
    # the step with and without the feedback, replayed from a CUDA graph
    sd_r, sd_c, _ = bench.synth_weights()
    for fb in (False, True):
        eng = RolloutEngine.from_state_dicts(sd_r, sd_c, dev)
        eng.set_graph({k: v.to(dev) for k, v in x.items()}, {k: v.to(dev) for k, v in ei.items()})
        if fb:
            eng.enable_geometry_feedback()
            eng.enable_event_selection()
        eng.capture(span=6, warmup=3)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            eng.step(6)
        e1.record()
        torch.cuda.synchronize()
        out['step_ms_with_feedback' if fb else 'step_ms'] = e0.elapsed_time(e1) / 20
        out['launches_with_feedback' if fb else 'launches'] = eng.launches_per_step
        if fb:
            out['finite'] = bool(torch.isfinite(eng.x['grain']).all() and torch.isfinite(eng.pred['edge_event']).all())
        del eng
    # host port in the reference's order (per-grain Python loop + numpy), bounded sample
    import grain_oracle as orc
    xs, eis, _, _ = lattice_domain((6, 6))
    t0 = time.perf_counter()
    orc.region_center(xs['joint'], eis[ET_GJ], xs['grain'].shape[0])
    dt = time.perf_counter() - t0
    out['cpu_port'] = {'grains': xs['grain'].shape[0], 'seconds': dt, 'grains_per_s': xs['grain'].shape[0] / dt,
                       'sample': '6x6 patches, oracle/grain_oracle.region_center (reference op order, 1 thread)'}
    out['gpu_grains_per_s'] = ng / (t_kernel * 1e-3)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, 'w') as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out))


if __name__ == '__main__':
    main()
