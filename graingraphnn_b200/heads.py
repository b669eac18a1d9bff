"""Output heads (kernel family (d)) and the in-place feature update (a12) as thin wrappers over the C ABI."""
import ctypes

import torch

from . import _lib
from ._lib import check, ptr
from .graph import _stream


def node_head(h, weight, bias, acts, area_in=None, area_scale=20.0, n_rows=None):
    """y = act(W h + b) per node (models.py:433-452). acts: per output row 0 none / 1 tanh / 2 relu.
    With `area_in` ([N] strided view of x_grain[:, 3]) also returns tanh(raw y0)/area_scale + area_in (models.py:445)."""
    n, C = h.shape
    n_out = weight.shape[0]
    y = torch.empty(n, n_out, dtype=torch.float32, device=h.device)
    area = torch.empty(n, dtype=torch.float32, device=h.device) if area_in is not None else None
    a = (ctypes.c_int32 * n_out)(*acts)
    with torch.cuda.device(h.device):
        check(_lib.lib().gg_node_head(ptr(h), h.stride(0), C, ptr(weight), ptr(bias), n_out, a, ptr(y), n_out,
                                      ptr(area_in), 0 if area_in is None else area_in.stride(0), float(area_scale),
                                      ptr(area), n if n_rows is None else n_rows, _stream()), 'gg_node_head')
    return y, area


def edge_head(h_joint, edge_index, edge_attr, w1, b1, w2, b2, want_edge=True):
    """edge_event = lin2([h[src], h[dst], a]), edge = tanh(lin1(...)) in original edge order (models.py:595-609)."""
    E = edge_index.shape[1]
    C = h_joint.shape[1]
    ev = torch.empty(E, dtype=torch.float32, device=h_joint.device)
    ed = torch.empty(E, 2, dtype=torch.float32, device=h_joint.device) if want_edge else None
    ei = edge_index.contiguous()
    ea = edge_attr.detach().float().reshape(-1).contiguous()
    with torch.cuda.device(h_joint.device):
        check(_lib.lib().gg_edge_head(ptr(h_joint), h_joint.stride(0), C, ptr(ei), E, ptr(ea), ptr(w1), ptr(b1),
                                      ptr(w2), ptr(b2), ptr(ev), ptr(ed), _stream()), 'gg_edge_head')
    return ev, ed


def feature_update(x_joint, x_grain, y_joint, y_grain, dz, z_max, scratch=None, n_joint=None, n_grain=None):
    """In place: models.py:510-516 + test.py:401-407.  x_* may be column-strided views of padded buffers."""
    if scratch is None:
        scratch = torch.empty(1, dtype=torch.int32, device=x_joint.device)
    assert x_joint.stride(1) == 1 and x_grain.stride(1) == 1
    with torch.cuda.device(x_joint.device):
        check(_lib.lib().gg_feature_update(ptr(x_joint), x_joint.stride(0), x_joint.shape[0] if n_joint is None else n_joint,
                                           ptr(y_joint), ptr(x_grain), x_grain.stride(0),
                                           x_grain.shape[0] if n_grain is None else n_grain, x_grain.shape[1], ptr(y_grain),
                                           float(dz), float(z_max), ptr(scratch), _stream()), 'gg_feature_update')


def feature_update_batched(x_joint, x_grain, y_joint, y_grain, dz_joint, dz_grain, z_max, n_joint=None, n_grain=None):
    """Ensemble form of feature_update: per-node z increments (span of the node's graph / 121), per-node clamp."""
    assert x_joint.stride(1) == 1 and x_grain.stride(1) == 1
    with torch.cuda.device(x_joint.device):
        check(_lib.lib().gg_feature_update_batched(ptr(x_joint), x_joint.stride(0), x_joint.shape[0] if n_joint is None else n_joint,
                                                   ptr(y_joint), ptr(x_grain), x_grain.stride(0),
                                                   x_grain.shape[0] if n_grain is None else n_grain, x_grain.shape[1], ptr(y_grain),
                                                   ptr(dz_joint), ptr(dz_grain), float(z_max), _stream()), 'gg_feature_update_batched')
