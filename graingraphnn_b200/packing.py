"""Weight packing: reference `state_dict` layout -> the fused buffers the kernels read.

All re-associations are linear and exact in real arithmetic (SURVEY.md §7 "Algebra"):
  * lin_key / lin_value / lin_query act per NODE on cat([X, h]); the periodic displacement of the source's x,y,z
    (periodGATconv.py:209-211) becomes a rank-3 correction with the first three weight columns (Wv3; for the key it
    collapses into the 4 scalars QX = [Wk[:, :3]^T q, We . q] per target and gate).
  * lin_l2, lin_edge and lin_skip move behind the aggregation (sum_e alpha_e = 1).
  * all gates of a cell share one projection GEMM; lin_skip of the edge types that end in the same node type are summed.
  * "raw scores" (cells without hidden state, i.e. the encoder, models.py:237-238): q_i . (Wk x_j) = x_j . (Wk^T q_i), so the
    key projection disappears altogether: the target carries Q'_i = [Wk[:, :F]^T q_i (F <= 15 source features) | We . q_i]
    (16 floats per gate instead of C + 4) and the source row only its raw features next to V.  With hidden state (decoder)
    the same identity holds on the full input [X padded to 32 | h]: the staged source row shrinks from K | V (2 G C floats) to
    [input (32 + C) | V (G C)] and Q' grows to 32 + C floats per gate (slot 31 = We . q) — fewer bytes per edge through the
    L2 -> shared-memory path that bounds the decoder gather.
Packing runs once per weight version (on whatever device the parameters live on) in float64, then rounds to fp32.
"""
import torch


def pad4(n):
    return (n + 3) // 4 * 4


class ConvWeights:
    """The tensors of one PeriodConv (periodGATconv.py:119-143), PyG [out, in] layout."""

    __slots__ = ('wk', 'bk', 'wq', 'bq', 'wv', 'bv', 'w2', 'b2', 'we', 'ws', 'bs')

    def __init__(self, conv):
        g = lambda lin, name: getattr(lin, name, None)  # noqa: E731
        self.wk, self.bk = conv.lin_key.weight, g(conv.lin_key, 'bias')
        self.wq, self.bq = conv.lin_query.weight, g(conv.lin_query, 'bias')
        self.wv, self.bv = conv.lin_value.weight, g(conv.lin_value, 'bias')
        self.w2, self.b2 = conv.lin_l2.weight, g(conv.lin_l2, 'bias')
        self.we = conv.lin_edge.weight
        self.ws, self.bs = conv.lin_skip.weight, g(conv.lin_skip, 'bias')

    def tensors(self):
        return [getattr(self, n) for n in self.__slots__ if getattr(self, n) is not None]

    def host(self):
        """The same weights as host tensors: packing is fp64 arithmetic on ~50 small matrices per cell, done on the CPU so
        that no fp64 ATen / cuBLAS kernels run on the device (only the packed fp32 buffers are uploaded)."""
        h = object.__new__(ConvWeights)
        for n in self.__slots__:
            v = getattr(self, n)
            setattr(h, n, None if v is None else v.detach().cpu())
        return h


def _cols(w, k1, k1p, k2):
    """[out, k1+k2] -> [out, k1p+k2] with zero columns inserted after the first k1 (a wider w is cut to its first k1+k2
    columns: the hidden-state columns of a cell that runs without hidden state)."""
    w = w.detach().double()[:, :k1 + k2]
    assert w.shape[1] == k1 + k2, (tuple(w.shape), k1, k2)
    out = torch.zeros(w.shape[0], k1p + k2, dtype=torch.float64, device=w.device)
    out[:, :k1] = w[:, :k1]
    out[:, k1p:] = w[:, k1:]
    return out


def _vec(b, n, dev):
    return torch.zeros(n, dtype=torch.float64, device=dev) if b is None else b.detach().double().reshape(n)


class PackedCell:
    """Packed weights of `len(gates)` HeteroConv{edge_type: PeriodConv} layers that read the same input.

    conv_of(gate, edge_type) -> ConvWeights ; gate_bias(gate, node_type) -> tensor [1,C] or None
    in_dims[node_type] = (K1, K2): widths of the two input pieces (X and h); K2 may be 0.
    """

    RAW_K = 16          # floats of the raw-feature block / of Q' per gate in raw-score mode

    def _pos_rows(self, e, off, k1p, k2, rows_w, rows_b):
        """Four identity columns (x, y, z, 0) of the TARGET's features right behind its Q | QX (or Q') block, so that the tiled
        gather stages a target's query and position (periodGATconv.py:209) with ONE bulk copy."""
        self.posoff[e] = off
        eye = torch.zeros(4, k1p + k2, dtype=torch.float64)
        eye[0, 0] = eye[1, 1] = eye[2, 2] = 1.0
        rows_w.append(eye.to(rows_w[-1].device)); rows_b.append(torch.zeros(4, dtype=torch.float64, device=rows_b[-1].device))
        return off + 4

    def __init__(self, edge_types, gates, in_dims, C, conv_of, gate_bias=None, weighted=True, device=None, raw_scores=False,
                 raw_hidden=False):
        self.edge_types, self.gates, self.C, self.G = list(edge_types), list(gates), C, len(gates)
        self.weighted = bool(weighted)
        self.raw_k, self.we_slot = 0, None
        if raw_scores and not raw_hidden:   # h == 0: only the feature columns of every weight act; the hidden columns are dropped here
            self.raw_k, self.we_slot = self.RAW_K, self.RAW_K - 1
            in_dims = {t: (k[0], 0) for t, k in in_dims.items()}
            assert all(pad4(k[0]) <= self.RAW_K - 1 for k in in_dims.values()), 'raw-score mode needs <= 15 (padded) features'
            self._full_k2 = C
        elif raw_scores:                    # raw scores on [X padded to 32 | h]: the staged source row is the cell input itself
            k2s = {k[1] for k in in_dims.values()}
            assert k2s == {C} and all(pad4(k[0]) <= 31 for k in in_dims.values()), 'raw scores with hidden state: K2 == C, <= 31 features'
            self.raw_k, self.we_slot = 32 + C, 31
        self.in_dims = dict(in_dims)
        self.node_types = list(self.in_dims)
        G = self.G
        GC = G * C
        self.k1p = {t: pad4(k[0]) for t, k in self.in_dims.items()}
        self.kin = {t: self.k1p[t] + self.in_dims[t][1] for t in self.in_dims}
        self.koff, self.voff, self.qoff, self.qxoff, self.posoff, self.ncols = {}, {}, {}, {}, {}, {}
        self.Wcat, self.bcat, self.Wskip, self.btot = {}, {}, {}, {}
        self.Wv3, self.W2, self.We, self.b2 = {}, {}, {}, {}
        self.into = {t: [e for e in self.edge_types if e[2] == t] for t in self.node_types}

        for t in self.node_types:
            k1, k2 = self.in_dims[t]
            k1p = self.k1p[t]
            rows_w, rows_b, off = [], [], 0
            for e in self.edge_types:              # source roles: K and V gate blocks
                if e[0] != t:
                    continue
                if self.raw_k:                     # raw-score mode: [raw input (16, or 32 + C with hidden state) | V gate block], one bulk copy per edge
                    self.koff[e] = off
                    dev0 = conv_of(self.gates[0], e).wv.device
                    eye = torch.zeros(self.raw_k, k1p + k2, dtype=torch.float64, device=dev0)
                    eye[:k1p, :k1p] = torch.eye(k1p, dtype=torch.float64, device=dev0)
                    if k2:                         # hidden state in slots 32 .. 32 + C
                        eye[32:32 + k2, k1p:] = torch.eye(k2, dtype=torch.float64, device=dev0)
                    rows_w.append(eye); rows_b.append(torch.zeros(self.raw_k, dtype=torch.float64, device=dev0))
                    off += self.raw_k
                    self.voff[e] = off
                    for g in self.gates:
                        cw = conv_of(g, e)
                        rows_w.append(_cols(cw.wv, k1, k1p, k2)); rows_b.append(_vec(cw.bv, C, cw.wv.device))
                    off += GC
                    continue
                for role, store in (('k', self.koff), ('v', self.voff)):
                    store[e] = off
                    for g in self.gates:
                        cw = conv_of(g, e)
                        w, b = (cw.wk, cw.bk) if role == 'k' else (cw.wv, cw.bv)
                        rows_w.append(_cols(w, k1, k1p, k2)); rows_b.append(_vec(b, C, w.device))
                    off += GC
            for e in self.edge_types:              # target roles: Q gate block, directly followed by its QX block
                if e[2] != t:                      # ([Wk[:, :3]^T q (3), We . q (1)] per gate) so one bulk copy stages both
                    continue
                if self.raw_k:                     # Q'[g] = Wk^T q laid out like the staged source row, We . q in the spare slot
                    self.qoff[e] = self.qxoff[e] = off
                    fs, hs = self.in_dims[e[0]]
                    for g in self.gates:
                        cw = conv_of(g, e)
                        wq, bq = _cols(cw.wq, k1, k1p, k2), _vec(cw.bq, C, cw.wq.device)
                        m = torch.zeros(C, self.raw_k, dtype=torch.float64, device=wq.device)
                        m[:, :fs] = cw.wk.detach().double()[:, :fs]
                        if hs:
                            m[:, 32:32 + hs] = cw.wk.detach().double()[:, fs:fs + hs]
                        m[:, self.we_slot] = cw.we.detach().double().reshape(C)
                        rw, rb = m.t() @ wq, m.t() @ bq
                        if g == self.gates[0]:     # the target's x, y, z ride in three spare slots of the first gate's Q': the source
                            p0 = self.we_slot - 3  # input is zero there (feature padding), so the score does not see them
                            assert pad4(fs) <= p0 and not rw[p0:p0 + 3].any() and not rb[p0:p0 + 3].any()
                            for i in range(3):
                                rw[p0 + i, i] = 1.0
                            self.posoff[e] = off + p0
                        rows_w.append(rw); rows_b.append(rb)
                    off += self.raw_k * G
                    continue
                self.qoff[e] = off
                for g in self.gates:
                    cw = conv_of(g, e)
                    rows_w.append(_cols(cw.wq, k1, k1p, k2)); rows_b.append(_vec(cw.bq, C, cw.wq.device))
                off += GC
                self.qxoff[e] = off
                for g in self.gates:
                    cw = conv_of(g, e)
                    wq, bq = _cols(cw.wq, k1, k1p, k2), _vec(cw.bq, C, cw.wq.device)
                    m = torch.cat([cw.wk.detach().double()[:, :3], cw.we.detach().double().reshape(C, 1)], dim=1)  # [C,4]
                    rows_w.append(m.t() @ wq); rows_b.append(m.t() @ bq)
                off += 4 * G
                off = self._pos_rows(e, off, k1p, k2, rows_w, rows_b)
            if off == 0:                            # a node type that is neither source nor target of anything
                rows_w.append(torch.zeros(4, k1p + k2, dtype=torch.float64)); rows_b.append(torch.zeros(4, dtype=torch.float64))
                off = 4
            self.ncols[t] = off
            self.Wcat[t] = torch.cat(rows_w, 0)
            self.bcat[t] = torch.cat(rows_b, 0)
            assert self.Wcat[t].shape == (off, k1p + k2)

            ws = torch.zeros(GC, k1p + k2, dtype=torch.float64, device=self.Wcat[t].device)
            bt = torch.zeros(GC, dtype=torch.float64, device=self.Wcat[t].device)
            for gi, g in enumerate(self.gates):
                for e in self.into[t]:
                    cw = conv_of(g, e)
                    ws[gi * C:(gi + 1) * C] += _cols(cw.ws, k1, k1p, k2)
                    bt[gi * C:(gi + 1) * C] += _vec(cw.bs, C, ws.device)
                if gate_bias is not None:
                    gb = gate_bias(g, t)
                    if gb is not None:
                        bt[gi * C:(gi + 1) * C] += gb.detach().double().reshape(C)
            self.Wskip[t], self.btot[t] = ws, bt

        for e in self.edge_types:
            wv3 = torch.zeros(GC, 4, dtype=torch.float64)
            w2, we, b2 = [], [], []
            for gi, g in enumerate(self.gates):
                cw = conv_of(g, e)
                wv3 = wv3.to(cw.wv.device)
                wv3[gi * C:(gi + 1) * C, :3] = cw.wv.detach().double()[:, :3]
                w2.append(cw.w2.detach().double()); we.append(cw.we.detach().double().reshape(C))
                b2.append(_vec(cw.b2, C, cw.w2.device))
            self.Wv3[e], self.W2[e] = wv3, torch.stack(w2, 0)
            self.We[e], self.b2[e] = torch.stack(we, 0), torch.stack(b2, 0)

        for d in (self.Wcat, self.bcat, self.Wskip, self.btot, self.Wv3, self.W2, self.We, self.b2):
            for k in d:
                d[k] = d[k].to(torch.float32).contiguous()
                if device is not None:
                    d[k] = d[k].to(device)

    def to(self, device):
        for d in (self.Wcat, self.bcat, self.Wskip, self.btot, self.Wv3, self.W2, self.We, self.b2):
            for k in d:
                d[k] = d[k].to(device)
        return self


def tf32_round(x):
    """cvt.rna.tf32.f32 on a float32 tensor: round-to-nearest (ties away) to 10 mantissa bits, low 13 bits cleared."""
    u = x.contiguous().view(torch.int32)
    return ((u + 0x1000) & -8192).view(torch.float32)


def split_tf32(w):
    """w = hi + lo with both parts exactly representable in TF32 (operands of the 3xTF32 tensor-core GEMM)."""
    hi = tf32_round(w)
    return hi, tf32_round(w - hi)


def tc_weight_layout(pk, t):
    """Wcat[t] re-laid out for gg_node_proj_tc: K = [features padded to 32 | hidden], split into (hi, lo)."""
    k1p, kin = pk.k1p[t], pk.kin[t]
    k2 = kin - k1p
    W = pk.Wcat[t]
    out = torch.zeros(W.shape[0], 32 + k2, dtype=torch.float32, device=W.device)
    out[:, :k1p] = W[:, :k1p]
    out[:, 32:] = W[:, k1p:]
    hi, lo = split_tf32(out)
    return hi.contiguous(), lo.contiguous()


def tc_gate_weight_layout(pk, t):
    """Wall[t] [G*C, Ktot] for gg_gate_update_tc: per gate row block, K = [lin_l2 of each incoming edge type (C each) |
    feature chunk (32) | summed lin_skip on the hidden state], split into (hi, lo).  The feature chunk carries the summed
    lin_skip columns of X and, in its tail (from RB = 32 - (2G+3)), the rank-1 terms the kernel feeds per node:
    [ea_0[g] -> lin_edge_0 | cnt_0 -> lin_l2 bias_0 | ea_1[g] -> lin_edge_1 | cnt_1 -> lin_l2 bias_1 | 1 -> btot]."""
    C, G = pk.C, pk.G
    k1p, kin = pk.k1p[t], pk.kin[t]
    k2 = kin - k1p
    ins = pk.into[t]
    rb = 32 - (2 * G + 3)
    assert len(ins) <= 2 and k1p <= rb, (len(ins), k1p, rb)
    ktot = len(ins) * C + 32 + k2
    dev = pk.Wskip[t].device
    W = torch.zeros(G * C, ktot, dtype=torch.float32, device=dev)
    for g in range(G):
        rows = slice(g * C, (g + 1) * C)
        for i, e in enumerate(ins):
            W[rows, i * C:(i + 1) * C] = pk.W2[e][g]
        off = len(ins) * C
        W[rows, off:off + k1p] = pk.Wskip[t][rows, :k1p]
        for i, e in enumerate(ins):
            W[rows, off + rb + i * (G + 1) + g] = pk.We[e][g]
            W[rows, off + rb + i * (G + 1) + G] = pk.b2[e][g]
        W[rows, off + 31] = pk.btot[t][rows]
        W[rows, off + 32:] = pk.Wskip[t][rows, k1p:]
    hi, lo = split_tf32(W)
    return hi.contiguous(), lo.contiguous(), ktot


def version_key(tensors):
    """Changes whenever any of the tensors is rebound, moved or modified in place (load_state_dict, .to(), optimizer)."""
    return tuple((t.data_ptr(), t._version, t.device.type, t.device.index) for t in tensors)
