"""`HeteroGCLSTM` / `HeteroGC` — the SAGEConv flavour of the recurrent cells, on the CUDA path.

Mirror of the reference cells (heterogclstm.py:21-196 and :199-275): same constructors, same
`forward(x_dict, edge_index_dict, h_dict, c_dict)` / `forward(x_dict, edge_index_dict)` signatures (no `edge_attr`),
same parameter tree (`conv_{i,f,c,o}.convs.<src>__<rel>__<dst>.{lin_l.weight, lin_l.bias, lin_r.weight}`,
the never-read `W_{i,f,c,o}.<type>` and the gate biases `b_{i,f,c,o}.<type>`).  The reference only reaches them with
`layers > 1`, where `SeqGCLSTM.forward` raises (SURVEY.md finding 1), so this is an API-complete module, not a tuned one.

Execution: per edge type one `gg_segment_mean` (PyG SAGEConv mean over cat([X, h]) of the sources, in two pieces so the
concat is never materialised), then per target type ONE `gg_gate_update` over Z = [mean_e0 | mean_e1 | ... | X] and h
whose weight rows are, per gate, [lin_l of every incoming edge type | sum of lin_r], fused with the LSTM / ReLU math.
"""
import torch
from torch import nn
from torch.nn import Parameter

from . import _lib
from ._lib import check, ptr
from .cell import _as_f32c, pad_features, require_cuda
from .graph import GLOBAL_CSR_CACHE, _stream
from .nn import HeteroConv, Linear, glorot_
from .packing import pad4, version_key


class SAGEConv(nn.Module):
    """Parameter holder for torch_geometric.nn.SAGEConv(in_channels=(-1,-1), aggr='mean', root_weight=True):
    `lin_l` ([C, K_src], bias) acts on the neighbour mean, `lin_r` ([C, K_dst], no bias) on the target row."""

    def __init__(self, in_channels, out_channels, bias=True, **kwargs):
        super().__init__()
        if kwargs.get('aggr', 'mean') != 'mean' or kwargs.get('normalize', False) or kwargs.get('project', False):
            raise NotImplementedError("SAGEConv: only aggr='mean' without normalize/project (the reference's setting)")
        if isinstance(in_channels, int):
            in_channels = (in_channels, in_channels)
        self.in_channels, self.out_channels = in_channels, out_channels
        self.lin_l = Linear(in_channels[0], out_channels, bias=bias)
        self.lin_r = Linear(in_channels[1], out_channels, bias=False)

    def materialize(self, k_src, k_dst):
        self.lin_l.materialize(k_src)
        self.lin_r.materialize(k_dst)

    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError('graingraphnn_b200 SAGEConv only stores parameters; the cells run the fused CUDA path')


class _SageBase(nn.Module):
    GATES = ()
    with_h = True

    def __init__(self, in_channels_dict, out_channels, metadata, bias=True, device='cpu'):
        super().__init__()
        if out_channels % 32 or not (32 <= out_channels <= 128):
            raise NotImplementedError('out_channels must be a multiple of 32 in [32, 128]')
        self.in_channels_dict, self.out_channels, self.metadata = in_channels_dict, out_channels, metadata
        self.bias, self.device = bias, device
        self._create_parameters_and_layers()
        self._pack = None

    def _make_conv(self, g):
        extra = self.out_channels if self.with_h else 0
        convs = {}
        for e in self.metadata[1]:
            conv = SAGEConv((-1, -1), self.out_channels, bias=self.bias)
            s, d = e[0], e[-1]
            if s in self.in_channels_dict and d in self.in_channels_dict:
                conv.materialize(self.in_channels_dict[s] + extra, self.in_channels_dict[d] + extra)
            convs[e] = conv
        setattr(self, f'conv_{g}', HeteroConv(convs))

    # -- packing: Wall[t] [G*C, Ktot_t], K layout [mean piece of every incoming edge type: (X_src pad4 | h_src) | X_t pad4 | h_t]
    def _packed(self, device):
        tensors = []
        for g in self.GATES:
            for e in self.metadata[1]:
                c = getattr(self, f'conv_{g}').conv(e)
                tensors += [c.lin_l.weight, c.lin_r.weight] + ([c.lin_l.bias] if c.lin_l.bias is not None else [])
            if hasattr(self, f'b_{g}'):
                tensors += list(getattr(self, f'b_{g}').values())
        key = (str(device), version_key(tensors))
        if self._pack is not None and self._pack[0] == key:
            return self._pack[1]
        C, G = self.out_channels, len(self.GATES)
        k2 = C if self.with_h else 0
        F = dict(self.in_channels_dict)
        Fp = {t: pad4(f) for t, f in F.items()}
        into = {t: [e for e in self.metadata[1] if e[-1] == t and e[0] in F] for t in F}
        pk = {'into': into, 'Fp': Fp, 'W': {}, 'b': {}, 'kz': {}, 'off': {}}
        for t in F:
            offs, off = {}, 0
            for e in into[t]:
                offs[e] = off
                off += Fp[e[0]] + k2
            kz = off + Fp[t]                               # columns of Z (means, then the node's own padded features)
            W = torch.zeros(G * C, kz + k2, dtype=torch.float64)
            b = torch.zeros(G * C, dtype=torch.float64)
            for gi, g in enumerate(self.GATES):
                rows = slice(gi * C, (gi + 1) * C)
                for e in into[t]:
                    conv = getattr(self, f'conv_{g}').conv(e)
                    wl, wr = conv.lin_l.weight.detach().double().cpu(), conv.lin_r.weight.detach().double().cpu()
                    fs = F[e[0]]
                    W[rows, offs[e]:offs[e] + fs] = wl[:, :fs]
                    W[rows, offs[e] + Fp[e[0]]:offs[e] + Fp[e[0]] + k2] = wl[:, fs:]
                    W[rows, off:off + F[t]] += wr[:, :F[t]]
                    W[rows, kz:kz + k2] += wr[:, F[t]:]
                    if conv.lin_l.bias is not None:
                        b[rows] += conv.lin_l.bias.detach().double().cpu()
                if hasattr(self, f'b_{g}'):
                    b[rows] += getattr(self, f'b_{g}')[t].detach().double().cpu().reshape(C)
            pk['W'][t], pk['b'][t] = W.float().contiguous().to(device), b.float().contiguous().to(device)
            pk['kz'][t], pk['off'][t] = kz, dict(offs, self_=off)
        self._pack = (key, pk)
        return pk

    def _run(self, x_dict, edge_index_dict, h, c, mode):
        for t, X in x_dict.items():
            require_cuda(X, f"x_dict['{t}']")
        L = _lib.lib()
        dev = next(iter(x_dict.values())).device
        pk = self._packed(dev)
        C, G = self.out_channels, len(self.GATES)
        k2 = C if self.with_h else 0
        xpad = {t: pad_features(x_dict[t], pk['Fp'][t]) for t in self.in_channels_dict}
        out_h, out_c = {}, {}
        lstm = mode == _lib.GG_GATE_LSTM
        with torch.cuda.device(dev):
            st = _stream()
            for t in self.in_channels_dict:
                ins = [e for e in pk['into'][t] if e in edge_index_dict]
                if not ins:
                    continue                                # PyG HeteroConv emits nothing for this node type
                if len(ins) != len(pk['into'][t]):
                    raise KeyError(f'edge_index_dict lacks an edge type that ends in {t!r}')
                n = xpad[t].shape[0]
                Z = torch.empty(n, pk['kz'][t], dtype=torch.float32, device=dev)
                for e in ins:
                    s = e[0]
                    csr = GLOBAL_CSR_CACHE.get(edge_index_dict[e], xpad[s].shape[0], n)
                    o = pk['off'][t][e]
                    check(L.gg_segment_mean(ptr(xpad[s]), xpad[s].stride(0), pk['Fp'][s], ptr(csr.rowptr), ptr(csr.col), n,
                                            Z.data_ptr() + 4 * o, Z.stride(0), st), 'gg_segment_mean')
                    if k2:
                        check(L.gg_segment_mean(ptr(h[s]), h[s].stride(0), C, ptr(csr.rowptr), ptr(csr.col), n,
                                                Z.data_ptr() + 4 * (o + pk['Fp'][s]), Z.stride(0), st), 'gg_segment_mean')
                Z[:, pk['off'][t]['self_']:] = xpad[t]
                out_h[t] = torch.empty(n, C if lstm else G * C, dtype=torch.float32, device=dev)
                if lstm:
                    out_c[t] = torch.empty(n, C, dtype=torch.float32, device=dev)
                W = pk['W'][t]
                check(L.gg_gate_update(None, 0, ptr(Z), Z.stride(0), pk['kz'][t], ptr(h[t]) if k2 else None, h[t].stride(0) if k2 else 0,
                                       ptr(W), W.stride(0), ptr(pk["b"][t]), ptr(c[t]) if (lstm and c is not None) else None,
                                       ptr(out_h[t]), ptr(out_c[t]) if lstm else None, n, G, C, mode, st), 'gg_gate_update')
        return out_h, out_c

    def __deepcopy__(self, memo):
        import copy
        new = self.__class__.__new__(self.__class__)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            new.__dict__[k] = None if k == '_pack' else copy.deepcopy(v, memo)
        return new


class HeteroGCLSTM(_SageBase):
    """LSTM cell whose gates are HeteroConv{SAGEConv} over cat([X, h]) (heterogclstm.py:125-196)."""
    GATES = ('i', 'f', 'c', 'o')

    def _create_parameters_and_layers(self):
        for g in self.GATES:                                # registration order conv_g, W_g, b_g (heterogclstm.py:51-89)
            self._make_conv(g)
            setattr(self, f'W_{g}', nn.ParameterDict({t: Parameter(torch.empty(k, self.out_channels))
                                                      for t, k in self.in_channels_dict.items()}))   # never read (:125-160)
            setattr(self, f'b_{g}', nn.ParameterDict({t: Parameter(torch.empty(1, self.out_channels))
                                                      for t in self.in_channels_dict}))
        for g in self.GATES:
            for p in list(getattr(self, f'W_{g}').values()) + list(getattr(self, f'b_{g}').values()):
                glorot_(p)

    def _set_hidden_state(self, x_dict, h_dict):
        if h_dict is None:
            h_dict = {t: torch.zeros(X.shape[0], self.out_channels, device=X.device) for t, X in x_dict.items()}
        return h_dict

    _set_cell_state = _set_hidden_state

    @torch.no_grad()
    def forward(self, x_dict, edge_index_dict, h_dict=None, c_dict=None):
        h = {t: _as_f32c(v) for t, v in self._set_hidden_state(x_dict, h_dict).items()}
        c = {t: _as_f32c(v) for t, v in self._set_cell_state(x_dict, c_dict).items()}
        return self._run(x_dict, edge_index_dict, h, c, _lib.GG_GATE_LSTM)


class HeteroGC(_SageBase):
    """relu(HeteroConv{SAGEConv}(x_dict)) — no hidden state, no gate bias (heterogclstm.py:236-275)."""
    GATES = ('i',)
    with_h = False

    def _create_parameters_and_layers(self):
        self._make_conv('i')

    @torch.no_grad()
    def forward(self, x_dict, edge_index_dict):
        out_h, _ = self._run(x_dict, edge_index_dict, None, None, _lib.GG_GATE_RELU)
        return out_h
