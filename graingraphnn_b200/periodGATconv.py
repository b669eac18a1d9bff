"""`PeriodConv` — periodic-aware dot-product-attention convolution, CUDA-only.

Mirror of the reference operator (periodGATconv.py:15-240): same constructor, same `forward(x, edge_index, edge_attr)`
signature, same parameter names/shapes (`lin_key/lin_query/lin_value/lin_l2/lin_edge/lin_skip`), so reference
checkpoints load unchanged.  The arithmetic of `message` + segment softmax + scatter-add (:204-236, :174-192) runs in
three hand-written sm_100a kernels (see cell.py); there is no CPU implementation.
"""
import torch
from torch import Tensor, nn

from . import _lib
from .cell import pad_features, require_cuda, run_cell
from .graph import GLOBAL_CSR_CACHE, permute_to_csr
from .nn import Linear
from .packing import ConvWeights, PackedCell, version_key


class PeriodConv(nn.Module):
    weighted = True   # periodconv.PeriodConv overrides with False (periodconv.py:235)

    def __init__(self, in_channels, out_channels, heads=1, concat=True, beta=False, dropout=0.,
                 edge_dim=None, bias=True, root_weight=True, **kwargs):
        super().__init__()
        if kwargs.get('aggr', 'add') not in ('add', 'sum'):
            raise NotImplementedError("PeriodConv aggregates with 'add' (periodGATconv.py:102)")
        if heads != 1:
            # the reference applies lin_l2 (H*C -> H*C) to a [E, H, C] view (:218), which only type-checks for H = 1
            raise NotImplementedError('PeriodConv supports heads=1 (the only setting the reference can run)')
        if beta and root_weight:
            raise NotImplementedError('beta-gated skip (lin_beta) is not used by GrainGNN and not implemented')
        if out_channels % 32 or not (32 <= out_channels <= 128):
            raise NotImplementedError('out_channels must be a multiple of 32 in [32, 128] (reference grid: 96/64/32)')
        edge_dim = 1                                     # periodGATconv.py:105 forces a scalar edge feature
        self.in_channels, self.out_channels, self.heads = in_channels, out_channels, heads
        self.beta, self.root_weight, self.concat = False, root_weight, concat
        self.dropout, self.edge_dim = dropout, edge_dim
        if isinstance(in_channels, int):
            in_channels = (in_channels, in_channels)
        self.lin_key = Linear(in_channels[0], heads * out_channels)
        self.lin_query = Linear(in_channels[1], heads * out_channels)
        self.lin_value = Linear(in_channels[0], heads * out_channels)
        self.lin_l2 = Linear(heads * out_channels, heads * out_channels, bias=bias)
        self.lin_edge = Linear(edge_dim, heads * out_channels, bias=False)
        self.lin_skip = Linear(in_channels[1], heads * out_channels, bias=bias)
        self.register_parameter('lin_beta', None)
        self._packed = None
        self._packed_key = None

    def reset_parameters(self):
        for lin in (self.lin_key, self.lin_query, self.lin_value, self.lin_l2, self.lin_edge, self.lin_skip):
            lin.reset_parameters()

    def materialize(self, d_src, d_dst):
        self.lin_key.materialize(d_src)
        self.lin_value.materialize(d_src)
        self.lin_query.materialize(d_dst)
        self.lin_skip.materialize(d_dst)

    def _pack(self, d_src, d_dst, same, device):
        cw = ConvWeights(self)
        key = (version_key(cw.tensors()), d_src, d_dst, same, str(device), self.root_weight)
        if self._packed is None or self._packed_key != key:
            et = ('n', 'e', 'n') if same else ('s', 'e', 'd')
            in_dims = {'n': (d_src, 0)} if same else {'s': (d_src, 0), 'd': (d_dst, 0)}
            hw = cw.host()
            pk = PackedCell([et], ['x'], in_dims, self.out_channels, lambda g, e: hw,
                            weighted=self.weighted, device=device)
            if not self.root_weight:
                for t in pk.Wskip:
                    pk.Wskip[t].zero_(); pk.btot[t].zero_()
            self._packed, self._packed_key = pk, key
        return self._packed

    def forward(self, x, edge_index, edge_attr=None, return_attention_weights=None):
        if return_attention_weights is not None:
            raise NotImplementedError('attention weights are never materialised by the fused kernel')
        if self.training and self.dropout > 0:
            raise NotImplementedError('attention dropout (training) is outside the inference path')
        assert edge_attr is not None                     # periodGATconv.py:221
        same = isinstance(x, Tensor)
        xs, xd = (x, x) if same else x
        require_cuda(xs, 'x'); require_cuda(edge_index, 'edge_index')
        self.materialize(xs.shape[1], xd.shape[1])
        pk = self._pack(xs.shape[1], xd.shape[1], same, xs.device)
        et = pk.edge_types[0]
        with torch.no_grad():
            csr = GLOBAL_CSR_CACHE.get(edge_index, xs.shape[0], xd.shape[0])
            ea = permute_to_csr(edge_attr.detach().float(), csr)
            if same:
                xpad = {'n': pad_features(xs, pk.k1p['n'])}
            else:
                xpad = {'s': pad_features(xs, pk.k1p['s']), 'd': pad_features(xd, pk.k1p['d'])}
            out, _ = run_cell(pk, xpad, None, None, {et: csr}, {et: ea}, _lib.GG_GATE_RAW)
        return out[et[2]]

    def __repr__(self):
        return f'{self.__class__.__name__}({self.in_channels}, {self.out_channels}, heads={self.heads})'
