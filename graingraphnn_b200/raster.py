"""Row f4: polygon raster and layer-error QoI on the device (graph_datastruct.py:553-610 `plot_polygons`, :346-348
`compute_error_layer`) — gg_raster_polygons / gg_count_mismatch (csrc/raster.cu).  No CPU fallback."""
import numpy as np
import torch

from . import _lib
from ._lib import check, ptr


def _stream():
    return torch.cuda.current_stream().cuda_stream


def plot_polygons(polygons, s, device='cuda'):
    """polygons: {grain id (1-based): [[x, y], ...]} in draw order (the reference's `region_coors`), coordinates in domain units;
    s = imagesize[0].  -> alpha_field int32 [s, s] on `device` (Image convention [ny, nx]); 0 where nothing was drawn."""
    device = torch.device(device)
    if device.type != 'cuda':
        raise RuntimeError('graingraphnn_b200 runs on CUDA devices only (no CPU fallback)')
    ids, ptrs, verts = [], [0], []
    for gid, poly in polygons.items():
        p = np.asarray(np.asarray(poly, dtype=np.float64) * s, dtype=int)        # :585 truncation toward zero
        if len(p) > 32:
            raise ValueError(f'grain {gid} has {len(p)} vertices (the raster kernel takes <= 32)')
        ids.append(int(gid))
        verts.append(p.reshape(-1, 2))
        ptrs.append(ptrs[-1] + len(p))
    n = len(ids)
    ids_t = torch.tensor(ids if n else [0], dtype=torch.int32, device=device)
    ptr_t = torch.tensor(ptrs, dtype=torch.int32, device=device)
    v = np.concatenate(verts, 0) if n else np.zeros((1, 2), dtype=int)
    verts_t = torch.from_numpy(np.ascontiguousarray(v, dtype=np.int32)).to(device)
    scratch = torch.empty(2 * s, 2 * s, dtype=torch.int32, device=device)
    alpha = torch.empty(s, s, dtype=torch.int32, device=device)
    with torch.cuda.device(device):
        check(_lib.lib().gg_raster_polygons(ptr(ptr_t), ptr(verts_t), ptr(ids_t), n, s, ptr(scratch), ptr(alpha), _stream()), 'gg_raster_polygons')
    return alpha


def error_layer(alpha_pde, alpha_field):
    """Fraction of pixels where the two int32 [s, s] fields differ (compute_error_layer, :346-348)."""
    a = alpha_field.contiguous()
    b = alpha_pde.to(a.device, torch.int32).contiguous()
    if a.shape != b.shape:
        raise ValueError((tuple(a.shape), tuple(b.shape)))
    count = torch.zeros(1, dtype=torch.int64, device=a.device)
    with torch.cuda.device(a.device):
        check(_lib.lib().gg_count_mismatch(ptr(a), ptr(b), a.numel(), ptr(count), _stream()), 'gg_count_mismatch')
    return float(count.item()) / a.numel()


def area_counts(alpha_field):
    """{grain id: pixels} of an alpha_field (graph_datastruct.py:287-288)."""
    c = torch.bincount(alpha_field.reshape(-1).long())
    nz = c.nonzero().view(-1)
    return dict(zip(nz.tolist(), c[nz].tolist()))
