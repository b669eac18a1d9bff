"""Small nn building blocks with the parameter names / shapes of the PyG modules the reference uses, so that
reference `state_dict`s load unchanged.  No arithmetic happens here: these classes only own
parameters; the cells pack them for the CUDA kernels.
"""
import math

import torch
from torch import nn
from torch.nn.parameter import Parameter, UninitializedParameter


class Linear(nn.Module):
    """Parameter holder equivalent to torch_geometric.nn.dense.linear.Linear as used at periodGATconv.py:119-131:
    `weight` is [out, in]; in_channels = -1 defers allocation until the input width is known (first forward or
    load_state_dict).  Default init = kaiming-uniform(a=sqrt 5) weight, U(+-1/sqrt(in)) bias."""

    def __init__(self, in_channels, out_channels, bias=True):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.weight = Parameter(torch.empty(out_channels, in_channels)) if in_channels > 0 else UninitializedParameter()
        if bias:
            self.bias = Parameter(torch.empty(out_channels))
        else:
            self.register_parameter('bias', None)
        self.reset_parameters()

    @property
    def materialized(self):
        return not isinstance(self.weight, UninitializedParameter)

    def reset_parameters(self):
        if self.in_channels <= 0:
            return
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        if self.bias is not None:
            bound = 1.0 / math.sqrt(self.in_channels)
            nn.init.uniform_(self.bias, -bound, bound)

    @torch.no_grad()
    def materialize(self, in_channels):
        if not self.materialized:
            self.in_channels = int(in_channels)
            self.weight.materialize((self.out_channels, self.in_channels))
            self.reset_parameters()
        elif self.in_channels != in_channels:
            raise ValueError(f'Linear expects {self.in_channels} input channels, got {in_channels}')

    def _load_from_state_dict(self, state_dict, prefix, *args, **kwargs):
        w = state_dict.get(prefix + 'weight')
        if w is not None and not self.materialized and not isinstance(w, UninitializedParameter):
            with torch.no_grad():
                self.in_channels = int(w.shape[-1])
                self.weight.materialize((self.out_channels, self.in_channels))
        super()._load_from_state_dict(state_dict, prefix, *args, **kwargs)

    def forward(self, x):  # pragma: no cover - the fused cells never call this
        raise RuntimeError('graingraphnn_b200.nn.Linear only stores parameters; arithmetic runs in the fused CUDA cells')

    def extra_repr(self):
        return f'{self.in_channels}, {self.out_channels}, bias={self.bias is not None}'


class HeteroConv(nn.Module):
    """Container with PyG HeteroConv's parameter naming: `convs['src__rel__dst']` (heteropgclstm.py:49-52)."""

    def __init__(self, convs, aggr='sum'):
        super().__init__()
        if aggr != 'sum':
            raise NotImplementedError("only aggr='sum' (the reference's setting) is supported")
        self.edge_types = list(convs.keys())
        self.convs = nn.ModuleDict({'__'.join(k): v for k, v in convs.items()})
        self.aggr = aggr

    def conv(self, edge_type):
        return self.convs['__'.join(edge_type)]


def glorot_(t):
    """torch_geometric.nn.inits.glorot (heteropgclstm.py:90-99)."""
    stdv = math.sqrt(6.0 / (t.size(-2) + t.size(-1)))
    with torch.no_grad():
        t.uniform_(-stdv, stdv)
