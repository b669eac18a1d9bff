from torch import Tensor

from torch_geometric.nn.conv.message_passing import MessagePassing
from torch_geometric.nn.dense.linear import Linear


class SAGEConv(MessagePassing):
    """PyG 2.1.0 SAGEConv defaults: lin_l(mean_j x_j) + lin_r(x_i); lin_r has no bias."""

    def __init__(self, in_channels, out_channels, aggr='mean', normalize=False, root_weight=True,
                 project=False, bias=True, **kwargs):
        super().__init__(aggr=aggr, node_dim=0, **kwargs)
        assert not normalize and not project
        self.in_channels, self.out_channels, self.root_weight = in_channels, out_channels, root_weight
        if isinstance(in_channels, int):
            in_channels = (in_channels, in_channels)
        self.lin_l = Linear(in_channels[0], out_channels, bias=bias)
        if root_weight:
            self.lin_r = Linear(in_channels[1], out_channels, bias=False)

    def forward(self, x, edge_index, size=None):
        if isinstance(x, Tensor):
            x = (x, x)
        out = self.propagate(edge_index, x=x, size=size)
        out = self.lin_l(out)
        if self.root_weight and x[1] is not None:
            out = out + self.lin_r(x[1])
        return out

    def message(self, x_j):
        return x_j
