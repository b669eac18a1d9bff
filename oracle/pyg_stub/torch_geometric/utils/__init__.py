import torch


def _seg_reduce(src, index, dim_size, reduce):
    shape = (dim_size,) + tuple(src.shape[1:])
    if reduce == 'sum':
        return torch.zeros(shape, dtype=src.dtype, device=src.device).index_add_(0, index, src)
    idx = index.view(-1, *([1] * (src.dim() - 1))).expand_as(src)
    if reduce == 'max':
        out = torch.full(shape, float('-inf'), dtype=src.dtype, device=src.device)
        out = out.scatter_reduce(0, idx, src, reduce='amax', include_self=True)
        return torch.where(torch.isinf(out), torch.zeros_like(out), out)  # torch-scatter fills empty with 0
    raise ValueError(reduce)


def softmax(src, index, ptr=None, num_nodes=None, dim=0):
    """PyG 2.1.0 utils.softmax: exp(src - max_seg) / (sum_seg + 1e-16)."""
    assert ptr is None and dim == 0
    N = int(index.max()) + 1 if num_nodes is None else num_nodes
    src_max = _seg_reduce(src, index, N, 'max').index_select(0, index)
    out = (src - src_max).exp()
    out_sum = _seg_reduce(out, index, N, 'sum').index_select(0, index)
    return out / (out_sum + 1e-16)
