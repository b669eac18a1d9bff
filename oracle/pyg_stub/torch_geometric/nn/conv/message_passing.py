import inspect

import torch
from torch import Tensor


class MessagePassing(torch.nn.Module):
    """PyG 2.1.0 MessagePassing restricted to: flow='source_to_target', dense [2,E] edge_index,
    aggr in {'add','sum','mean'}, no fused message_and_aggregate, identity update."""

    def __init__(self, aggr='add', flow='source_to_target', node_dim=-2, **kwargs):
        super().__init__()
        assert flow == 'source_to_target'
        self.aggr, self.node_dim = aggr, node_dim
        self._msg_params = [p for p in inspect.signature(self.message).parameters]

    def propagate(self, edge_index, size=None, **kwargs):
        assert isinstance(edge_index, Tensor) and edge_index.dim() == 2 and edge_index.size(0) == 2
        assert self.node_dim == 0
        src_idx, dst_idx = edge_index[0], edge_index[1]
        size = [None, None] if size is None else list(size)
        out_kwargs = {}
        for name in self._msg_params:
            if name.endswith('_i') or name.endswith('_j'):
                data = kwargs.get(name[:-2], None)
                side = 1 if name.endswith('_i') else 0
                if isinstance(data, (tuple, list)):
                    assert len(data) == 2
                    for s in (0, 1):
                        if isinstance(data[s], Tensor) and size[s] is None:
                            size[s] = data[s].size(0)
                    data = data[side]
                elif isinstance(data, Tensor):
                    for s in (0, 1):
                        if size[s] is None:
                            size[s] = data.size(0)
                if isinstance(data, Tensor):
                    data = data.index_select(0, dst_idx if side == 1 else src_idx)
                out_kwargs[name] = data
        if size[1] is None:
            size[1] = size[0]
        special = {'index': dst_idx, 'ptr': None, 'size_i': size[1], 'size_j': size[0],
                   'edge_index': edge_index, 'dim_size': size[1]}
        for name in self._msg_params:
            if name in out_kwargs:
                continue
            if name in special:
                out_kwargs[name] = special[name]
            else:
                out_kwargs[name] = kwargs.get(name, None)
        msg = self.message(**out_kwargs)
        return self.aggregate(msg, dst_idx, None, size[1])

    def message(self, x_j):
        return x_j

    def aggregate(self, inputs, index, ptr=None, dim_size=None):
        shape = (dim_size,) + tuple(inputs.shape[1:])
        out = torch.zeros(shape, dtype=inputs.dtype, device=inputs.device).index_add_(0, index, inputs)
        if self.aggr in ('add', 'sum'):
            return out
        if self.aggr == 'mean':
            cnt = torch.zeros(dim_size, dtype=inputs.dtype, device=inputs.device).index_add_(
                0, index, torch.ones_like(index, dtype=inputs.dtype)).clamp_(min=1)
            return out / cnt.view(-1, *([1] * (inputs.dim() - 1)))
        raise ValueError(self.aggr)
