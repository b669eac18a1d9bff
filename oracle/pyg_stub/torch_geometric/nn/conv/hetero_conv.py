from collections import defaultdict

import torch
from torch.nn import Module, ModuleDict


def group(xs, aggr):
    if len(xs) == 0:
        return None
    if aggr is None:
        return torch.stack(xs, dim=1)
    if len(xs) == 1:
        return xs[0]
    out = torch.stack(xs, dim=0)
    out = getattr(torch, aggr)(out, dim=0)
    return out[0] if isinstance(out, tuple) else out


class HeteroConv(Module):
    """PyG 2.1.0 HeteroConv: ModuleDict keyed '__'.join(edge_type); iterates edge_index_dict order."""

    def __init__(self, convs, aggr='sum'):
        super().__init__()
        self.convs = ModuleDict({'__'.join(k): v for k, v in convs.items()})
        self.aggr = aggr

    def forward(self, x_dict, edge_index_dict, *args_dict, **kwargs_dict):
        out_dict = defaultdict(list)
        for edge_type, edge_index in edge_index_dict.items():
            src, rel, dst = edge_type
            str_edge_type = '__'.join(edge_type)
            if str_edge_type not in self.convs:
                continue
            args = []
            for value_dict in args_dict:
                if edge_type in value_dict:
                    args.append(value_dict[edge_type])
                elif src == dst and src in value_dict:
                    args.append(value_dict[src])
                elif src in value_dict or dst in value_dict:
                    args.append((value_dict.get(src, None), value_dict.get(dst, None)))
            kwargs = {}
            for arg, value_dict in kwargs_dict.items():
                arg = arg[:-5]
                if edge_type in value_dict:
                    kwargs[arg] = value_dict[edge_type]
                elif src == dst and src in value_dict:
                    kwargs[arg] = value_dict[src]
                elif src in value_dict or dst in value_dict:
                    kwargs[arg] = (value_dict.get(src, None), value_dict.get(dst, None))
            conv = self.convs[str_edge_type]
            if src == dst:
                out = conv(x_dict[src], edge_index, *args, **kwargs)
            else:
                out = conv((x_dict[src], x_dict[dst]), edge_index, *args, **kwargs)
            out_dict[dst].append(out)
        for key, value in out_dict.items():
            out_dict[key] = group(value, self.aggr)
        return out_dict
