#!/usr/bin/env python
"""Where do the warps of the tiled gather spend their cycles?  Needs a library built with GG_TILED_PROFILE=1:
    GG_TILED_PROFILE=1 python -m graingraphnn_b200.build && python scripts/gather_profile.py
Prints, per gather launch of one rollout step, the mean cycles per consumer / producer warp spent waiting on the stage
barriers and working."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'oracle'))
import bench  # noqa: E402
from graingraphnn_b200 import _lib  # noqa: E402
from graingraphnn_b200.engine import RolloutEngine  # noqa: E402


def main():
    dev = torch.device('cuda:0')
    x, ei, ea, _, _ = bench.make_domain(bench.weak_lxd(1))
    sd_r, sd_c, _ = bench.synth_weights()
    eng = RolloutEngine.from_state_dicts(sd_r, sd_c, device=dev)
    eng.set_graph({k: v.to(dev) for k, v in x.items()}, {k: v.to(dev) for k, v in ei.items()}, {k: v.to(dev) for k, v in ea.items()})
    for _ in range(2):
        eng.step(6)
    L = _lib.lib()
    L.gg_gather_tiled_profile.restype = ctypes.c_int
    buf = (ctypes.c_ulonglong * 8)()
    real = L.gg_pgat_gather_tiled_multi_multi
    rows = []

    def wrapped(*a):
        L.gg_gather_tiled_profile(buf)            # reset
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = real(*a)
        e1.record()
        assert L.gg_gather_tiled_profile(buf) == 0, 'library was not built with GG_TILED_PROFILE=1'
        v = list(buf)
        rows.append((e0.elapsed_time(e1) * 1e3, v[0] / max(v[4], 1), v[1] / max(v[4], 1), v[2] / max(v[5], 1), v[3] / max(v[5], 1), v[6] * 1e3 / max(v[7], 1)))
        return rc

    class Proxy:
        def __getattr__(self, name):
            return wrapped if name == 'gg_pgat_gather_tiled_multi' else getattr(L, name)

    _lib._LIB = Proxy()
    eng.step(6)
    torch.cuda.synchronize()
    _lib._LIB = L
    print('launch     us | consumer wait   busy (kcycles/warp) | producer wait   busy | SM MHz during the launch')
    for i, r in enumerate(rows):
        print(f'{i:6d} {r[0]:7.1f} | {r[1] / 1e3:10.1f} {r[2] / 1e3:8.1f} | {r[3] / 1e3:10.1f} {r[4] / 1e3:8.1f} | {r[5]:7.0f}')


if __name__ == '__main__':
    main()
