"""`HeteroPGCLSTM` / `HeteroPGC` — recurrent heterogeneous graph cells on the fused CUDA path.

Mirror of the reference cells (heteropgclstm.py:18-183 and :185-284): same constructors, same
`forward(x_dict, edge_index_dict, edge_attr, h_dict, c_dict) -> (h_dict, c_dict)`, same parameter tree
(`conv_{i,f,c,o}.convs.<src>__<rel>__<dst>.lin_*`, `b_{i,f,c,o}.<node_type>`), so `regressor0.pt` /
`classifier1.pt` load unchanged through `load_state_dict`.

What differs is the execution: the 4 gates x 3 edge types = 12 PeriodConv calls of the reference (≈300 op dispatches)
become 2 projection GEMMs + 3 fused gather kernels + 2 gate-GEMM/LSTM kernels (cell.py).  When both h and c are absent
(the encoder call, models.py:237-238) the forget gate is dead (f * 0) and only (i, c, o) are evaluated on the raw
feature columns.
"""
import torch
from torch import nn
from torch.nn import Parameter

from . import _lib
from .cell import _as_f32c, pad_features, raw_gather_available, raw_hidden_available, require_cuda, run_cell
from .graph import GLOBAL_CSR_CACHE, permute_to_csr
from .nn import HeteroConv, glorot_
from .packing import ConvWeights, PackedCell, version_key
from .periodGATconv import PeriodConv


class _PGCBase(nn.Module):
    GATES = ()
    conv_class = PeriodConv

    def __init__(self, in_channels_dict, out_channels, metadata, bias=True, device='cpu'):
        super().__init__()
        self.in_channels_dict = in_channels_dict
        self.out_channels = out_channels
        self.metadata = metadata
        self.bias = bias
        self.device = device
        self._create_parameters_and_layers()
        self._set_parameters()
        self._packs = {}

    # -- parameters: registration order conv_g then b_g, like heteropgclstm.py:48-82 -------------------------
    def _make_gate(self, g):
        convs = {}
        for edge_type in self.metadata[1]:
            conv = self.conv_class(in_channels=(-1, -1), out_channels=self.out_channels, bias=self.bias)
            s, d = edge_type[0], edge_type[-1]
            if s in self.in_channels_dict and d in self.in_channels_dict:   # widths are known: cat([X, h])
                conv.materialize(self.in_channels_dict[s] + self.out_channels,
                                 self.in_channels_dict[d] + self.out_channels)
            convs[edge_type] = conv
        setattr(self, f'conv_{g}', HeteroConv(convs))
        setattr(self, f'b_{g}', nn.ParameterDict({t: Parameter(torch.empty(1, self.out_channels))
                                                  for t in self.in_channels_dict}))

    def _create_parameters_and_layers(self):
        for g in self.GATES:
            self._make_gate(g)

    def _set_parameters(self):
        for g in self.GATES:
            for key in getattr(self, f'b_{g}'):
                glorot_(getattr(self, f'b_{g}')[key])

    def _set_hidden_state(self, x_dict, h_dict):
        if h_dict is None:
            h_dict = {t: torch.zeros(X.shape[0], self.out_channels, device=X.device) for t, X in x_dict.items()}
        return h_dict

    def _set_cell_state(self, x_dict, c_dict):
        if c_dict is None:
            c_dict = {t: torch.zeros(X.shape[0], self.out_channels, device=X.device) for t, X in x_dict.items()}
        return c_dict

    # -- packing ---------------------------------------------------------------------------------------------
    def _weights(self, gates):
        return {(g, e): ConvWeights(getattr(self, f'conv_{g}').conv(e)) for g in gates for e in self.metadata[1]}

    def packed(self, gates, with_h, device, edge_types=None, raw=None):
        """raw: pack for the raw-score gather (no key projection; cells without hidden state only).  None = whenever the fast
        gather path will take it (sm_100 device, <= 4 gates, <= 15 padded features per node type)."""
        edge_types = tuple(self.metadata[1]) if edge_types is None else tuple(edge_types)
        if raw is None:
            raw = len(gates) <= 4 and str(device).startswith('cuda') and \
                (raw_hidden_available(len(gates), self.out_channels) if with_h else raw_gather_available()) \
                and all((f + 3) // 4 * 4 <= PackedCell.RAW_K - 1 for f in self.in_channels_dict.values())
        cws = self._weights(gates)
        tensors = [t for cw in cws.values() for t in cw.tensors()]
        tensors += [getattr(self, f'b_{g}')[t] for g in gates for t in self.in_channels_dict]
        key = (tuple(gates), with_h, str(device), edge_types, bool(raw), version_key(tensors))
        slot = (tuple(gates), with_h, edge_types, bool(raw))
        hit = self._packs.get(slot)
        if hit is None or hit[0] != key:
            C = self.out_channels
            in_dims = {t: (f, C) for t, f in self.in_channels_dict.items()}
            hws = {k: cw.host() for k, cw in cws.items()}             # pack on the host (fp64), upload the fp32 result
            pk = PackedCell(edge_types, gates, in_dims, C, lambda g, e: hws[(g, e)],
                            gate_bias=lambda g, t: getattr(self, f'b_{g}')[t].detach().cpu(),
                            weighted=self.conv_class.weighted, device=device, raw_scores=bool(raw), raw_hidden=bool(raw) and with_h)
            if not with_h and not raw:   # h == 0: only the feature columns of every weight matter (K = K1p)
                for t in pk.node_types:
                    k1p = pk.k1p[t]
                    pk.Wcat[t] = pk.Wcat[t][:, :k1p].contiguous()
                    pk.Wskip[t] = pk.Wskip[t][:, :k1p].contiguous()
                    pk.kin[t] = k1p
            self._packs[slot] = (key, pk)
            hit = self._packs[slot]
        return hit[1]

    def _prepare(self, x_dict, edge_index_dict, edge_attr):
        for t, X in x_dict.items():
            require_cuda(X, f"x_dict['{t}']")
        if edge_attr is None:
            raise ValueError('edge_attr (dict of [E,1] edge lengths) is required: PeriodConv asserts it (periodGATconv.py:221)')
        edge_types = [e for e in edge_index_dict if e in self.metadata[1]]
        csr, ea = {}, {}
        for e in edge_types:
            s, d = e[0], e[-1]
            csr[e] = GLOBAL_CSR_CACHE.get(edge_index_dict[e], x_dict[s].shape[0], x_dict[d].shape[0])
            ea[e] = permute_to_csr(_as_f32c(edge_attr[e]), csr[e])
        return edge_types, csr, ea

    def __deepcopy__(self, memo):
        import copy
        cls = self.__class__
        new = cls.__new__(cls)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            new.__dict__[k] = {} if k == '_packs' else copy.deepcopy(v, memo)
        return new


class HeteroPGCLSTM(_PGCBase):
    """LSTM cell whose gates are HeteroConv{PeriodConv} over cat([X, h]) (heteropgclstm.py:111-183)."""
    GATES = ('i', 'f', 'c', 'o')

    @torch.no_grad()
    def forward(self, x_dict, edge_index_dict, edge_attr=None, h_dict=None, c_dict=None):
        edge_types, csr, ea = self._prepare(x_dict, edge_index_dict, edge_attr)
        dev = next(iter(x_dict.values())).device
        fresh = h_dict is None and c_dict is None
        if fresh:
            pk = self.packed(('i', 'c', 'o'), False, dev, edge_types)
            mode, h, c = _lib.GG_GATE_LSTM0, None, None
        else:
            pk = self.packed(self.GATES, True, dev, edge_types)
            mode = _lib.GG_GATE_LSTM
            h = {t: _as_f32c(v) for t, v in self._set_hidden_state(x_dict, h_dict).items()}
            c = None if c_dict is None else {t: _as_f32c(v) for t, v in c_dict.items()}
        xpad = {t: pad_features(x_dict[t], pk.k1p[t]) for t in pk.node_types}
        out_h, out_c = run_cell(pk, xpad, h, c, csr, ea, mode)
        return out_h, out_c


class HeteroPGC(_PGCBase):
    """Single ReLU graph-convolution layer, relu(conv_i(cat([X, h])) + b_i); c passes through (heteropgclstm.py:243-284)."""
    GATES = ('i',)

    @torch.no_grad()
    def forward(self, x_dict, edge_index_dict, edge_attr=None, h_dict=None, c_dict=None):
        edge_types, csr, ea = self._prepare(x_dict, edge_index_dict, edge_attr)
        dev = next(iter(x_dict.values())).device
        c_dict = self._set_cell_state(x_dict, c_dict)
        if h_dict is None:
            pk, h = self.packed(self.GATES, False, dev, edge_types), None
        else:
            pk = self.packed(self.GATES, True, dev, edge_types)
            h = {t: _as_f32c(v) for t, v in h_dict.items()}
        xpad = {t: pad_features(x_dict[t], pk.k1p[t]) for t in pk.node_types}
        out_h, _ = run_cell(pk, xpad, h, None, csr, ea, _lib.GG_GATE_RELU)
        return out_h, c_dict
