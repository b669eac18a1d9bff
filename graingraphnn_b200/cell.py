"""Execution of one packed cell (all gates of a HeteroConv{PeriodConv} stack) on the CUDA kernels.

Three kernel families per call, in this order:
  (c)  gg_node_proj   per node type   : P_t = [X_t | h_t] Wcat_t^T + bcat_t
  (b)  gg_pgat_gather per edge type   : agg_e, ea_e  (periodic wrap + scores + segment softmax + aggregation)
  (c') gg_gate_update per node type   : lin_l2 / lin_edge / lin_skip + biases + gate non-linearities / LSTM update
"""
import ctypes
import os

import torch
import torch.nn.functional as F

from . import _lib
from ._lib import AggInput, GatherSegment, check, ptr
from .graph import _stream, edge_wrap


def require_cuda(t, what):
    if not t.is_cuda:
        raise RuntimeError(f'graingraphnn_b200: {what} must live on a CUDA device; this package has no CPU path')


def pad_features(x, k1p):
    """[N, K1] -> contiguous fp32 [N, K1p] (zero columns appended) — rows become 16-byte aligned."""
    x = x.detach()
    if x.dtype != torch.float32:
        x = x.float()
    if x.shape[1] == k1p:
        return x.contiguous()
    return F.pad(x, (0, k1p - x.shape[1])).contiguous()


def gemm_mode():
    """'tc' (tcgen05 3xTF32 projections) or 'simt' (fp32 CUDA cores); GG_GEMM overrides, default = tc when supported."""
    m = os.environ.get('GG_GEMM', 'auto').lower()
    if m in ('tc', 'simt'):
        return m
    global _AUTO
    if _AUTO is None:
        L = _lib.lib()
        _AUTO = 'tc' if hasattr(L, 'gg_tc_supported') and L.gg_tc_supported() == 1 else 'simt'
    return _AUTO


_AUTO = None


def raw_gather_available():
    """True when gg_pgat_gather will run its bulk-copy kernel, the only one that implements raw-score mode."""
    if os.environ.get('GG_GATHER', '').lower().startswith('l'):
        return False
    L = _lib.lib()
    return hasattr(L, 'gg_tc_supported') and L.gg_tc_supported() == 1


def raw_hidden_available(G, C):
    """True when cells WITH hidden state may be packed for raw scores (source row = [input | V]): only the warp-specialised
    gather implements that form.  GG_RAWH=0 keeps the K | V form (A/B measurements)."""
    mode = os.environ.get('GG_GATHER', '').lower()
    if mode.startswith('l') or mode.startswith('i') or os.environ.get('GG_RAWH', '1') == '0':
        return False
    L = _lib.lib()
    if not (hasattr(L, 'gg_tc_supported') and L.gg_tc_supported() == 1):
        return False
    return int(L.gg_gather_tile_ecap(G, C, 32 + C)) > 0


def tiled_ecap(pk, e):
    """Tile size of the warp-specialised gather for edge type e of this pack, 0 when that kernel does not apply (unweighted
    sum variant, G > 4, operand blocks not adjacent, not an sm_100 device, GG_GATHER=ldg|items)."""
    mode = os.environ.get('GG_GATHER', '').lower()
    if mode.startswith('l') or mode.startswith('i') or not pk.weighted or pk.G > 4 or pk.C % 32:
        return 0
    if pk.voff[e] != pk.koff[e] + (pk.raw_k if pk.raw_k else pk.G * pk.C):
        return 0
    if not pk.raw_k and pk.qxoff[e] != pk.qoff[e] + pk.G * pk.C:
        return 0
    if pk.posoff.get(e) != (pk.qoff[e] + pk.we_slot - 3 if pk.raw_k else pk.qxoff[e] + 4 * pk.G):
        return 0                                   # the target's position: right behind Q | QX, or in the spare slots of the first Q'
    L = _lib.lib()
    if not (hasattr(L, 'gg_tc_supported') and L.gg_tc_supported() == 1):
        return 0
    return int(L.gg_gather_tile_ecap(pk.G, pk.C, pk.raw_k))


def run_cell(pk, xpad, h, c, csr, ea_csr, mode, out_h=None, out_c=None, work=None, n_rows=None, wrap=None):
    """xpad[t]: [N_t, K1p] fp32 (x,y,z in columns 0..2); h[t]: [N_t, K2] or None; c[t]: [N_t, C] or None;
    csr[e]: EdgeCSR; ea_csr[e]: [E] edge attribute in CSR order.  Returns (out_h, out_c) dicts.
    n_rows[t] (optional): number of leading rows of node type t that results are computed for (the owned rows of a slab
    partition; the rows behind them are halo copies that only serve as message sources).
    wrap[e] (optional): per-edge wrap codes of gg_edge_wrap for the positions in xpad; computed here when absent."""
    L = _lib.lib()
    C, G = pk.C, pk.G
    GC = G * C
    st = _stream()
    work = {} if work is None else work
    dev = next(iter(xpad.values())).device

    def buf(name, shape):
        """flat grow-only workspace per name, viewed at the requested shape (encoder/decoder share storage)"""
        numel = 1
        for d in shape:
            numel *= int(d)
        b = work.get(name)
        if b is None or b.numel() < numel or b.device != dev:
            b = torch.empty(max(numel, 1), dtype=torch.float32, device=dev)
            work[name] = b
        return b[:numel].view(shape)

    with torch.cuda.device(dev):
        # (c) node projections
        P = {}
        use_tc = gemm_mode() == 'tc' and pk.C % 32 == 0 and max(pk.k1p.values()) <= 32
        for t in pk.node_types:
            x = xpad[t]
            n = x.shape[0]
            ht = None if h is None else h[t]
            P[t] = buf(('P', t), (n, pk.ncols[t]))
            if use_tc:
                k2 = 0 if ht is None else ht.shape[1]
                kp = 32 + k2
                if not hasattr(pk, 'tcW'):
                    pk.tcW = {}
                if t not in pk.tcW:
                    from .packing import tc_weight_layout
                    pk.tcW[t] = tc_weight_layout(pk, t)
                whi, wlo = pk.tcW[t]
                if os.environ.get('GG_PROJ', 'fused') != 'split' and hasattr(L, 'gg_node_proj_fused') and k2 % 32 == 0:
                    check(L.gg_node_proj_fused(ptr(x), x.stride(0), pk.k1p[t], ptr(ht), 0 if ht is None else ht.stride(0), k2,
                                               ptr(whi), ptr(wlo), pk.ncols[t], ptr(pk.bcat[t]), ptr(P[t]), pk.ncols[t], n, 0, st),
                          'gg_node_proj_fused')
                    continue
                ahi, alo = buf(('Ahi', t), (n, kp)), buf(('Alo', t), (n, kp))
                check(L.gg_split_tf32(ptr(x), x.stride(0), pk.k1p[t], ptr(ht), 0 if ht is None else ht.stride(0), k2,
                                      n, ptr(ahi), ptr(alo), kp, 32, st), 'gg_split_tf32')
                check(L.gg_node_proj_tc(ptr(ahi), ptr(alo), kp, pk.k1p[t], ptr(whi), ptr(wlo), pk.ncols[t], ptr(pk.bcat[t]),
                                        ptr(P[t]), pk.ncols[t], n, 0, st), 'gg_node_proj_tc')
                continue
            check(L.gg_node_proj(ptr(x), x.stride(0), pk.k1p[t],
                                 ptr(ht), 0 if ht is None else ht.stride(0), 0 if ht is None else ht.shape[1],
                                 ptr(pk.Wcat[t]), pk.kin[t], ptr(pk.bcat[t]),
                                 ptr(P[t]), pk.ncols[t], n, pk.ncols[t], st), 'gg_node_proj')
        # (b) fused gather: the edge types the warp-specialised kernel serves go into ONE launch per cell (GG_GATHER_MERGE=0: one
        # launch per edge type), the others through the item-list kernel
        agg, ea = {}, {}
        merged = []
        for e in pk.edge_types:
            s, _, d = e
            nd = xpad[d].shape[0]
            nd_out = nd if n_rows is None else n_rows[d]
            agg[e] = buf(('agg', e), (nd, GC))
            ea[e] = buf(('ea', e), (nd, G))
            g = csr[e]
            wr = wrap[e] if wrap is not None else edge_wrap(g, xpad[s], xpad[d])
            ecap = tiled_ecap(pk, e)
            if ecap:
                tiles, cta_ptr, n_ctas = g.tiles(ecap)      # builds nz / nzptr on first use
                seg = GatherSegment(P[s].data_ptr(), pk.ncols[s], pk.koff[e], P[d].data_ptr(), pk.ncols[d], pk.qoff[e],
                                    g.rowptr.data_ptr(), g.col.data_ptr() if g.n_edges else None, ea_csr[e].data_ptr() if g.n_edges else None,
                                    wr.data_ptr(), g.nz.data_ptr(), g.nzptr.data_ptr(), tiles.data_ptr(), cta_ptr.data_ptr(),
                                    g.n_edges, pk.Wv3[e].data_ptr(), nd_out, agg[e].data_ptr(), GC, ea[e].data_ptr())
                merged.append((seg, n_ctas, ecap, wr))      # wr stays referenced until the launch that reads it is enqueued
                continue
            check(L.gg_pgat_gather(ptr(P[s]), pk.ncols[s], pk.koff[e], pk.voff[e],
                                   ptr(P[d]), pk.ncols[d], pk.qoff[e], pk.qxoff[e],
                                   ptr(xpad[s]), xpad[s].stride(0), ptr(xpad[d]), xpad[d].stride(0),
                                   ptr(g.rowptr), ptr(g.col), ptr(ea_csr[e]), ptr(g.items), ptr(g.item_ptr), ptr(wr), pk.raw_k, ptr(pk.Wv3[e]),
                                   nd_out, G, C, 1 if pk.weighted else 0, ptr(agg[e]), GC, ptr(ea[e]), st), 'gg_pgat_gather')
        if merged:
            one = os.environ.get('GG_GATHER_MERGE', '1') == '0'
            groups = [[m] for m in merged] if one else [merged[i:i + 3] for i in range(0, len(merged), 3)]
            for grp in groups:
                arr = (GatherSegment * len(grp))(*[m[0] for m in grp])
                check(L.gg_pgat_gather_tiled_multi(arr, len(grp), grp[0][1], grp[0][2], pk.raw_k, G, C, st), 'gg_pgat_gather_tiled_multi')
        # (c') gate GEMM + LSTM update per node type
        out_h = {} if out_h is None else out_h
        out_c = {} if out_c is None else out_c
        lstm = mode in (_lib.GG_GATE_LSTM, _lib.GG_GATE_LSTM0)
        for t in pk.node_types:
            x = xpad[t]
            n = x.shape[0]
            n_out = n if n_rows is None else n_rows[t]
            ins = pk.into[t]
            if not ins:          # PyG HeteroConv emits nothing for a node type no edge type ends in
                continue
            arr = (AggInput * len(ins))()
            for i, e in enumerate(ins):
                arr[i] = AggInput(agg[e].data_ptr(), GC, ea[e].data_ptr(), csr[e].rowptr.data_ptr(),
                                  pk.W2[e].data_ptr(), pk.We[e].data_ptr(), pk.b2[e].data_ptr(), 1 if pk.weighted else 0)
            ht = None if h is None else h[t]
            ct = None if c is None else c[t]
            if t not in out_h:
                out_h[t] = torch.empty((n, C if lstm else GC), dtype=torch.float32, device=dev)
            if lstm and t not in out_c:
                out_c[t] = torch.empty((n, C), dtype=torch.float32, device=dev)
            if use_tc and 1 <= len(ins) <= 2 and (lstm or G == 1) and pk.k1p[t] <= 32 - (2 * G + 3):     # multi-gate RAW output stays on the fp32 SIMT kernel
                if not hasattr(pk, 'tcG'):
                    pk.tcG = {}
                if t not in pk.tcG:
                    from .packing import tc_gate_weight_layout
                    pk.tcG[t] = tc_gate_weight_layout(pk, t)
                ghi, glo, ktot = pk.tcG[t]
                check(L.gg_gate_update_tc(arr, len(ins), ptr(x), x.stride(0), pk.k1p[t],
                                          ptr(ht), 0 if ht is None else ht.stride(0),
                                          ptr(ghi), ptr(glo), ktot, ptr(ct), ptr(out_h[t]),
                                          ptr(out_c[t]) if lstm else None, n_out, G, C, mode, 0, st), 'gg_gate_update_tc')
                continue
            check(L.gg_gate_update(arr, len(ins), ptr(x), x.stride(0), pk.k1p[t],
                                   ptr(ht), 0 if ht is None else ht.stride(0),
                                   ptr(pk.Wskip[t]), pk.kin[t], ptr(pk.btot[t]),
                                   ptr(ct), ptr(out_h[t]), ptr(out_c[t]) if lstm else None,
                                   n_out, G, C, mode, st), 'gg_gate_update')
    return out_h, out_c


def _as_f32c(t):
    t = t.detach()
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()
