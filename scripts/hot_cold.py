#!/usr/bin/env python
"""Does sustained load slow the memory-bound launches?  In-step gather times and a plain 1 GiB device copy, measured (a) after 3
warm-up steps, (b) right after 80 back-to-back steps, (c) after 2 s of idle; nvidia-smi clocks (SM / memory) and power alongside."""
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from graingraphnn_b200 import _lib, engine as _engine  # noqa: E402
from graingraphnn_b200.engine import RolloutEngine  # noqa: E402


def main():
    dev = torch.device('cuda:0')
    x, ei, ea, glob, _ = bench.make_domain(bench.weak_lxd(1))
    sd_r, sd_c, _ = bench.synth_weights()
    eng = RolloutEngine.from_state_dicts(sd_r, sd_c, device=dev)
    eng.set_graph({k: v.to(dev) for k, v in x.items()}, {k: v.to(dev) for k, v in ei.items()}, {k: v.to(dev) for k, v in ea.items()}, global_pos=glob)
    _engine._TWO_STREAMS = False
    L = _lib.lib()
    real = L.gg_pgat_gather_tiled_multi
    a = torch.empty(1 << 29, dtype=torch.bfloat16, device=dev)
    b = torch.empty_like(a)
    rows = []
    proc = subprocess.Popen(['nvidia-smi', '--query-gpu=clocks.sm,clocks.mem,power.draw,clocks_event_reasons.sw_power_cap', '--format=csv,noheader,nounits', '-lms', '20'],
                            stdout=subprocess.PIPE, text=True)
    threading.Thread(target=lambda: [rows.append((time.time(), l.strip())) for l in proc.stdout], daemon=True).start()

    def measure(tag):
        evs = []

        def wrapped(*args):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); rc = real(*args); e1.record()
            evs.append((e0, e1))
            return rc

        class Proxy:
            def __getattr__(self, name):
                return wrapped if name == 'gg_pgat_gather_tiled_multi' else getattr(L, name)
        _lib._LIB = Proxy()
        t0 = time.time()
        torch.cuda._sleep(40_000_000)
        eng.step(6)
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record(); b.copy_(a); c1.record()
        torch.cuda.synchronize()
        t1 = time.time()
        _lib._LIB = L
        g = [x0.elapsed_time(x1) * 1e3 for x0, x1 in evs]
        smi = [r[1] for r in rows if t0 <= r[0] <= t1]
        print(f'{tag:34s} gather {[round(v) for v in g]} sum {sum(g) / 1e3:.3f} ms   copy {2 * a.numel() * 2 / (c0.elapsed_time(c1) * 1e-3) / 1e9:.0f} GB/s   smi {smi[len(smi) // 2] if smi else None}', flush=True)

    for _ in range(3):
        eng.step(6)
    torch.cuda.synchronize()
    measure('after 3 warm-up steps')
    for rep in range(2):
        for _ in range(80):
            eng.step(6)
        measure(f'right after 80 eager steps ({rep})')
    eng.capture(6, warmup=1)
    for _ in range(150):
        eng.step(6)
    measure('right after 150 graph replays')
    torch.cuda.synchronize()
    time.sleep(2.0)
    measure('after 2 s idle')
    proc.terminate()


if __name__ == '__main__':
    main()
