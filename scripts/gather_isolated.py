#!/usr/bin/env python
"""Per-launch time of the gather family, in the step and replayed alone (L2 flushed before each replay), on the generate-mode
domain and on the regular honeycomb stand-in, engine rows in caller or Morton order:
    python scripts/gather_isolated.py [generated|honeycomb] [morton|caller]
Separates what the graph costs (isolated times) from what the neighbouring kernels of the step cost (in-step minus isolated)."""
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
    graph = sys.argv[1] if len(sys.argv) > 1 else 'generated'
    order = sys.argv[2] if len(sys.argv) > 2 else 'morton'
    dev = torch.device('cuda:0')
    if graph == 'generated':
        x, ei, ea, glob, _ = bench.make_domain(bench.weak_lxd(1))
    else:
        from graingraphnn_b200.synth import lattice_domain
        x, ei, glob, _ = lattice_domain((36, 30))
        glob = {t: v[:, :2] for t, v in glob.items()} if isinstance(glob, dict) else glob
        ea = None
    sd_r, sd_c, _ = bench.synth_weights()
    eng = RolloutEngine.from_state_dicts(sd_r, sd_c, device=dev)
    eng.set_graph({k: v.to(dev) for k, v in x.items()}, {k: v.to(dev) for k, v in ei.items()},
                  None if ea is None else {k: v.to(dev) for k, v in ea.items()}, global_pos=glob if order == 'morton' else None)
    _engine._TWO_STREAMS = False
    os.environ['GG_GATHER_MERGE'] = os.environ.get('GG_GATHER_MERGE', '0')
    for _ in range(3):
        eng.step(6)
    L = _lib.lib()
    real = L.gg_pgat_gather_tiled_multi
    calls, evs = [], []

    def wrapped(arr, n, *rest):
        keep = (ctypes.c_byte * ctypes.sizeof(arr)).from_buffer_copy(arr)          # the segment structs of this launch
        calls.append((keep, n, rest))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = real(arr, n, *rest)
        e1.record()
        evs.append((e0, e1))
        return rc

    class Proxy:
        def __getattr__(self, name):
            return wrapped if name == 'gg_pgat_gather_tiled_multi' else getattr(L, name)

    _lib._LIB = Proxy()
    torch.cuda._sleep(40_000_000)
    eng.step(6)
    torch.cuda.synchronize()
    _lib._LIB = L
    in_step = [a.elapsed_time(b) * 1e3 for a, b in evs]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    iso = []
    for keep, n, rest in calls:
        arr = ctypes.cast(keep, ctypes.POINTER(_lib.GatherSegment))
        ts = []
        for _ in range(5):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            real(arr, n, *rest)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        iso.append(sorted(ts)[len(ts) // 2])
    c = eng.counts()
    print(f'{graph} rows={order} grains={c["n_grain"]} merge={os.environ["GG_GATHER_MERGE"]}')
    print('in step :', [round(v) for v in in_step], 'sum %.3f ms' % (sum(in_step) / 1e3))
    print('isolated:', [round(v) for v in iso], 'sum %.3f ms' % (sum(iso) / 1e3))


if __name__ == '__main__':
    main()
