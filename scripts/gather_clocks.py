#!/usr/bin/env python
"""SM clock during every gather launch of eager rollout steps (is the launch that follows a tensor-core GEMM slower because the
power cap pulls the clock down?).  Needs the profile build of the library:
    python -c "from graingraphnn_b200 import build as b; b.build_variant('prof', defines=['GG_TILED_PROFILE'])"
    GG_LIB=$PWD/graingraphnn_b200/lib/variants/prof.so python scripts/gather_clocks.py [steps]
Prints per launch position of a step: mean event time, mean SM MHz (clock64 / globaltimer inside the kernel) over the steps."""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from graingraphnn_b200 import _lib, engine as _engine  # noqa: E402
from graingraphnn_b200.engine import RolloutEngine  # noqa: E402


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    dev = torch.device('cuda:0')
    x, ei, ea, _, _ = bench.make_domain(bench.weak_lxd(1))
    sd_r, sd_c, _ = bench.synth_weights()
    eng = RolloutEngine.from_state_dicts(sd_r, sd_c, device=dev)
    eng.set_graph({k: v.to(dev) for k, v in x.items()}, {k: v.to(dev) for k, v in ei.items()}, {k: v.to(dev) for k, v in ea.items()})
    _engine._TWO_STREAMS = False
    for _ in range(3):
        eng.step(6)
    L = _lib.lib()
    ring = (ctypes.c_ulonglong * 256)()
    L.gg_gather_tiled_clocks.restype = ctypes.c_int
    assert L.gg_gather_tiled_clocks(ring) >= 0, 'library was not built with GG_TILED_PROFILE'
    real = L.gg_pgat_gather_tiled_multi_multi
    evs = []

    def wrapped(*a):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = real(*a)
        e1.record()
        evs.append((e0, e1))
        return rc

    class Proxy:
        def __getattr__(self, name):
            return wrapped if name == 'gg_pgat_gather_tiled_multi' else getattr(L, name)

    _lib._LIB = Proxy()
    torch.cuda._sleep(40_000_000)
    for _ in range(steps):
        eng.step(6)
    torch.cuda.synchronize()
    _lib._LIB = L
    n = L.gg_gather_tiled_clocks(ring)
    per = len(evs) // steps
    assert n == len(evs) == per * steps, (n, len(evs))
    print('pos      us    MHz   (mean over %d back-to-back eager steps; positions 0-5 encoder R, C; 6-11 decoder R, C)' % steps)
    tot = 0.0
    for pos in range(per):
        us = [evs[s * per + pos][0].elapsed_time(evs[s * per + pos][1]) * 1e3 for s in range(steps)]
        mhz = [ring[2 * ((s * per + pos) & 127)] * 1e3 / max(ring[2 * ((s * per + pos) & 127) + 1], 1) for s in range(steps)]
        tot += sum(us) / steps
        print(f'{pos:3d} {sum(us) / steps:7.1f} {sum(mhz) / steps:6.0f}   first step {us[0]:6.1f} us {mhz[0]:5.0f} MHz')
    print('gather per step: %.3f ms' % (tot / 1e3))


if __name__ == '__main__':
    main()
