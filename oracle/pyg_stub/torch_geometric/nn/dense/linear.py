import copy
import math

import torch
import torch.nn.functional as F
from torch import nn
from torch.nn.parameter import Parameter, UninitializedParameter


class Linear(nn.Module):
    """PyG 2.1.0 nn.dense.linear.Linear: weight [out, in]; in_channels=-1 is lazy."""

    def __init__(self, in_channels, out_channels, bias=True, weight_initializer=None, bias_initializer=None):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        if in_channels > 0:
            self.weight = Parameter(torch.empty(out_channels, in_channels))
        else:
            self.weight = UninitializedParameter()
            self._hook = self.register_forward_pre_hook(self.initialize_parameters)
        if bias:
            self.bias = Parameter(torch.empty(out_channels))
        else:
            self.register_parameter('bias', None)
        self._load_hook = self._register_load_state_dict_pre_hook(self._lazy_load_hook)
        self.reset_parameters()

    def __deepcopy__(self, memo):
        out = Linear(self.in_channels, self.out_channels, self.bias is not None)
        if self.in_channels > 0:
            out.weight = copy.deepcopy(self.weight, memo)
        if self.bias is not None:
            out.bias = copy.deepcopy(self.bias, memo)
        return out

    def reset_parameters(self):
        if self.in_channels > 0:
            nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
            if self.bias is not None:
                bound = 1.0 / math.sqrt(self.in_channels)
                nn.init.uniform_(self.bias, -bound, bound)
        elif self.bias is not None and self.in_channels > 0:
            pass

    def forward(self, x):
        return F.linear(x, self.weight, self.bias)

    @torch.no_grad()
    def initialize_parameters(self, module, input):
        if isinstance(self.weight, UninitializedParameter):
            self.in_channels = input[0].size(-1)
            self.weight.materialize((self.out_channels, self.in_channels))
            self.reset_parameters()
        self._hook.remove()
        delattr(self, '_hook')

    def _lazy_load_hook(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs):
        weight = state_dict.get(prefix + 'weight', None)
        if weight is not None and isinstance(self.weight, UninitializedParameter) \
                and not isinstance(weight, UninitializedParameter):
            self.in_channels = weight.size(-1)
            self.weight.materialize((self.out_channels, self.in_channels))
            if hasattr(self, '_hook'):
                self._hook.remove()
                delattr(self, '_hook')
