"""Torch restatement of what each CUDA kernel computes FROM THE PACKED BUFFERS (test-side only).

Lets the CPU suite check the host logic — weight packing, column layout, the algebraic re-association — against the
oracle without a GPU.  The kernels themselves are checked against the oracle on the GPU (tests/test_gpu_parity.py).
"""
import math

import torch

from graingraphnn_b200 import _lib


def emu_node_proj(x, h, W, b):
    a = x if h is None else torch.cat([x, h], 1)
    return a @ W[:, :a.shape[1]].t() + b


def emu_gather(pk, e, P_src, P_dst, pos_src, pos_dst, rowptr, col, ea_csr):
    C, G = pk.C, pk.G
    nd = P_dst.shape[0]
    agg = torch.zeros(nd, G * C, dtype=P_src.dtype)
    ea_out = torch.zeros(nd, G, dtype=P_src.dtype)
    dst = torch.repeat_interleave(torch.arange(nd), (rowptr[1:] - rowptr[:-1]).long())
    src = col.long()
    w = pos_src[src, :3] - pos_dst[dst, :3]
    w = (w < -0.5).to(P_src.dtype) - (w > 0.5).to(P_src.dtype)
    t = w - pos_dst[dst, :3]
    rk = getattr(pk, 'raw_k', 0)
    for g in range(G):
        V = P_src[src, pk.voff[e] + g * C: pk.voff[e] + (g + 1) * C]
        if rk:      # raw-score mode: the source row carries its raw features, the target Q' = [Wk^T q | We . q] (16 per gate)
            xj = P_src[src, pk.koff[e]: pk.koff[e] + rk].clone()
            xj[:, pk.we_slot] = ea_csr
            Qp = P_dst[dst, pk.qoff[e] + rk * g: pk.qoff[e] + rk * (g + 1)]
            raw_s = (Qp * xj).sum(1) + (Qp[:, :3] * w).sum(1)
        else:
            K = P_src[src, pk.koff[e] + g * C: pk.koff[e] + (g + 1) * C]
            Q = P_dst[dst, pk.qoff[e] + g * C: pk.qoff[e] + (g + 1) * C]
            QX = P_dst[dst, pk.qxoff[e] + 4 * g: pk.qxoff[e] + 4 * g + 4]
            raw_s = (Q * K).sum(1) + (QX[:, :3] * w).sum(1) + QX[:, 3] * ea_csr
        if pk.weighted:
            s = raw_s / math.sqrt(C)
            m = torch.full((nd,), float('-inf'), dtype=s.dtype).scatter_reduce(0, dst, s, 'amax')
            p = (s - m[dst]).exp()
            den = torch.zeros(nd, dtype=s.dtype).index_add_(0, dst, p)
            alpha = p / (den[dst] + 1e-16)
        else:
            alpha = torch.ones(src.shape[0], dtype=P_src.dtype)
        v = torch.relu(V + t @ pk.Wv3[e][g * C:(g + 1) * C, :3].to(P_src.dtype).t())
        agg[:, g * C:(g + 1) * C].index_add_(0, dst, v * alpha[:, None])
        ea_out[:, g].index_add_(0, dst, alpha * ea_csr)
    return agg, ea_out


def emu_gate_update(pk, t, agg, ea, rowptr, x, h, c, mode):
    C, G = pk.C, pk.G
    dt = x.dtype
    a = x if h is None else torch.cat([x, h], 1)
    pre = a @ pk.Wskip[t].to(dt)[:, :a.shape[1]].t() + pk.btot[t].to(dt)
    for e in pk.into[t]:
        deg = (rowptr[e][1:] - rowptr[e][:-1]).to(dt)
        cnt = (deg > 0).to(dt) if pk.weighted else deg
        for g in range(G):
            pre[:, g * C:(g + 1) * C] += agg[e][:, g * C:(g + 1) * C] @ pk.W2[e][g].to(dt).t() \
                + ea[e][:, g:g + 1] * pk.We[e][g].to(dt) + cnt[:, None] * pk.b2[e][g].to(dt)
    if mode == _lib.GG_GATE_RAW:
        return pre, None
    if mode == _lib.GG_GATE_RELU:
        return torch.relu(pre), None
    if mode == _lib.GG_GATE_LSTM:
        i, f, cc, o = (pre[:, k * C:(k + 1) * C] for k in range(4))
        c0 = torch.zeros_like(i) if c is None else c
        c2 = torch.sigmoid(f) * c0 + torch.sigmoid(i) * torch.tanh(cc)
    else:
        i, cc, o = (pre[:, k * C:(k + 1) * C] for k in range(3))
        c2 = torch.sigmoid(i) * torch.tanh(cc)
    return torch.sigmoid(o) * torch.tanh(c2), c2


def emu_cell(pk, xpad, h, c, csr, ea_csr, mode):
    """csr[e] = (rowptr, col) int tensors; everything on CPU in the dtype of xpad."""
    dt = next(iter(xpad.values())).dtype
    P = {t: emu_node_proj(xpad[t], None if h is None else h[t], pk.Wcat[t].to(dt), pk.bcat[t].to(dt)) for t in pk.node_types}
    agg, ea = {}, {}
    for e in pk.edge_types:
        s, _, d = e
        agg[e], ea[e] = emu_gather(pk, e, P[s], P[d], xpad[s], xpad[d], csr[e][0], csr[e][1], ea_csr[e])
    out_h, out_c = {}, {}
    rp = {e: csr[e][0] for e in pk.edge_types}
    for t in pk.node_types:
        if not pk.into[t]:
            continue
        out_h[t], out_c[t] = emu_gate_update(pk, t, agg, ea, rp, xpad[t], None if h is None else h[t],
                                             None if c is None else c[t], mode)
    return out_h, out_c
