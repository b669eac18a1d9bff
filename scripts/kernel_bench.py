#!/usr/bin/env python
"""Micro-benchmarks of the tensor-core kernels through the C ABI (device timing, CUDA events):
    python scripts/kernel_bench.py gate|proj [--sms 148,111,74,37]
Used to separate per-SM limits (time ~ 1/SMs) from chip-wide limits (L2 / HBM: time saturates)."""
import argparse
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from graingraphnn_b200 import _lib  # noqa: E402
from graingraphnn_b200._lib import AggInput, check, ptr  # noqa: E402
from graingraphnn_b200.packing import split_tf32  # noqa: E402


def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


def bench_gate(sms, M=249120, G=4, C=96, n_in=2, has_h=True):
    L, d = _lib.lib(), torch.device('cuda')
    st = torch.cuda.current_stream().cuda_stream
    agg = [torch.randn(M, G * C, device=d) for _ in range(n_in)]
    ea = [torch.rand(M, G, device=d) for _ in range(n_in)]
    rowptr = torch.arange(0, 3 * (M + 1), 3, dtype=torch.int32, device=d)
    We, b2 = torch.randn(G, C, device=d), torch.randn(G, C, device=d)
    X, H = torch.rand(M, 8, device=d), (torch.randn(M, C, device=d) if has_h else None)
    ktot = n_in * C + 32 + (C if has_h else 0)
    whi, wlo = split_tf32(torch.randn(G * C, ktot, device=d) * 0.05)
    btot, cin = torch.randn(G * C, device=d), torch.randn(M, C, device=d)
    oh, oc = torch.empty(M, C, device=d), torch.empty(M, C, device=d)
    arr = (AggInput * n_in)()
    for i in range(n_in):
        arr[i] = AggInput(agg[i].data_ptr(), G * C, ea[i].data_ptr(), rowptr.data_ptr(), 0, We.data_ptr(), b2.data_ptr(), 1)
    mode = _lib.GG_GATE_LSTM if G == 4 else _lib.GG_GATE_LSTM0
    for n in sms:
        us = timeit(lambda: check(L.gg_gate_update_tc(arr, n_in, ptr(X), 8, 8, ptr(H), C if has_h else 0, ptr(whi), ptr(wlo), ktot,
                                                      ptr(cin), ptr(oh), ptr(oc), M, G, C, mode, n, st), 'gg_gate_update_tc'))
        stages = ((M + 127) // 128) * G * (n_in * C // 32 + 1 + (C // 32 if has_h else 0))
        kb = 16 + 2 * C * 128 / 1024
        print(f'gate M={M} G={G} n_in={n_in} h={has_h} sms={n:3d}: {us:8.1f} us   {stages * kb * 1024 / us / 1e6:6.2f} TB/s L2->smem '
              f'({stages * kb * 1024 / us / 1e3 / n / 1.965:5.1f} B/clk/SM)  {2 * M * G * C * ktot * 3 / us / 1e6:7.1f} TF/s tf32 issued')


def bench_proj(sms, M=249120, N=2048, K2=96):
    L, d = _lib.lib(), torch.device('cuda')
    st = torch.cuda.current_stream().cuda_stream
    kp = 32 + K2
    ahi, alo = split_tf32(torch.randn(M, kp, device=d))
    whi, wlo = split_tf32(torch.randn(N, kp, device=d) * 0.05)
    bias, out = torch.randn(N, device=d), torch.empty(M, N, device=d)
    for n in sms:
        us = timeit(lambda: check(L.gg_node_proj_tc(ptr(ahi), ptr(alo), kp, 8, ptr(whi), ptr(wlo), N, ptr(bias), ptr(out), N, M, n, st), 'gg_node_proj_tc'))
        print(f'proj M={M} N={N} Kp={kp} sms={n:3d}: {us:8.1f} us   out {M * N * 4 / us / 1e6:5.2f} TB/s   {2 * M * N * kp * 3 / us / 1e6:7.1f} TF/s tf32 issued')
        x, h = torch.rand(M, 8, device=d), (torch.randn(M, K2, device=d) if K2 else None)
        us = timeit(lambda: check(L.gg_node_proj_fused(ptr(x), 8, 8, ptr(h), K2, K2, ptr(whi), ptr(wlo), N, ptr(bias), ptr(out), N, M, n, st), 'gg_node_proj_fused'))
        print(f'proj (fused split) sms={n:3d}: {us:8.1f} us   out {M * N * 4 / us / 1e6:5.2f} TB/s')


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('what', choices=['gate', 'proj', 'all'])
    ap.add_argument('--sms', default='148,111,74,37')
    a = ap.parse_args()
    sms = [int(v) for v in a.sms.split(',')]
    if a.what in ('gate', 'all'):
        bench_gate(sms)
        bench_gate(sms[:1], G=3, has_h=False)
        bench_gate(sms[:1], M=124560, n_in=1)
    if a.what in ('proj', 'all'):
        bench_proj(sms)
        bench_proj(sms[:1], N=704, K2=0)
