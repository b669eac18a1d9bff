"""dst-sorted CSR index of one edge type (kernel family (a)) and a cache keyed on tensor identity.

The reference rebuilds nothing: PyG's `propagate` (called at periodGATconv.py:174) re-gathers from the COO
`edge_index` on every one of the 48 PeriodConv calls of a rollout step.  Here the CSR is built once per edge_index
tensor and reused by all gates / cells / models until the topology update rebinds a new tensor
(models.py:840-841) or edits it in place (detected through the tensor's version counter).
"""
import weakref

import torch

from . import _lib
from ._lib import check, ptr


def _stream():
    return torch.cuda.current_stream().cuda_stream


class EdgeCSR:
    """rowptr[n_dst+1], col[E] (source ids), perm[E] (original edge ids) — int32, on the edge_index device;
    items[*, 4] / item_ptr[n_dst+1]: flat work list of the gather kernel (gg_csr_items)."""

    __slots__ = ('rowptr', 'col', 'perm', 'n_src', 'n_dst', 'n_edges', 'items', 'item_ptr', 'nz', 'nzptr', 'nz_count', '_tiles')

    def __init__(self, rowptr, col, perm, n_src, n_dst, n_edges, items=None, item_ptr=None):
        self.rowptr, self.col, self.perm = rowptr, col, perm
        self.n_src, self.n_dst, self.n_edges = n_src, n_dst, n_edges
        self.items, self.item_ptr = items, item_ptr
        self.nz = self.nzptr = self.nz_count = None
        self._tiles = {}

    def tiles(self, ecap):
        """Tile index of the warp-specialised gather (gg_csr_compact + gg_csr_tiles) for tiles of `ecap` in-edges: (tiles [*, 4],
        cta_ptr [n_ctas + 1], n_ctas), built on first use per tile size, then reused until the topology changes."""
        t = self._tiles.get(ecap)
        if t is None:
            L = _lib.lib()
            dev = self.rowptr.device
            with torch.cuda.device(dev):
                if self.nz is None:
                    self.nz = torch.empty(max(self.n_dst, 1), dtype=torch.int32, device=dev)
                    self.nzptr = torch.empty(self.n_dst + 1, dtype=torch.int32, device=dev)
                    self.nz_count = torch.empty(1, dtype=torch.int32, device=dev)
                    scratch = torch.empty(self.n_dst + 1, dtype=torch.int32, device=dev)
                    ws_bytes = L.gg_csr_workspace_bytes(0, self.n_dst)
                    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
                    check(L.gg_csr_compact(ptr(self.rowptr), self.n_dst, ptr(self.nz), ptr(self.nzptr), ptr(self.nz_count),
                                           ptr(scratch), ptr(ws), ws_bytes, _stream()), 'gg_csr_compact')
                n_ctas = L.gg_gather_ctas()
                tl = torch.empty(max(L.gg_csr_tiles_capacity(self.n_edges, ecap, n_ctas), 1), 4, dtype=torch.int32, device=dev)
                cta_ptr = torch.empty(n_ctas + 1, dtype=torch.int32, device=dev)
                scr = torch.empty(L.gg_csr_tiles_scratch_ints(self.n_edges, ecap, n_ctas), dtype=torch.int32, device=dev)
                check(L.gg_csr_tiles(ptr(self.nzptr), ptr(self.nz_count), self.n_edges, ecap, n_ctas, ptr(tl), ptr(cta_ptr), ptr(scr),
                                     _stream()), 'gg_csr_tiles')
                t = (tl, cta_ptr, n_ctas)
            self._tiles[ecap] = t
        return t


def build_csr(edge_index, n_src, n_dst, validate=True):
    """Stable counting sort of `edge_index` ([2,E] int64, CUDA) by target. Raises on out-of-range endpoints."""
    if not edge_index.is_cuda:
        raise RuntimeError('graingraphnn_b200 runs on CUDA tensors only (no CPU fallback)')
    if edge_index.dtype != torch.int64 or edge_index.dim() != 2 or edge_index.shape[0] != 2:
        raise ValueError('edge_index must be an int64 tensor of shape [2, E]')
    ei = edge_index.contiguous()
    E = ei.shape[1]
    dev = ei.device
    L = _lib.lib()
    rowptr = torch.empty(n_dst + 1, dtype=torch.int32, device=dev)
    col = torch.empty(E, dtype=torch.int32, device=dev)
    perm = torch.empty(E, dtype=torch.int32, device=dev)
    status = torch.empty(1, dtype=torch.int32, device=dev)
    ws_bytes = L.gg_csr_workspace_bytes(E, n_dst)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        check(L.gg_csr_build(ptr(ei), E, n_src, n_dst, ptr(rowptr), ptr(col), ptr(perm), ptr(status),
                             ptr(ws), ws_bytes, _stream()), 'gg_csr_build')
    # flat work list of the gather kernel: <= n_dst + E / dcap items of <= dcap consecutive in-edges
    dcap = L.gg_gather_dcap()
    item_ptr = torch.empty(n_dst + 1, dtype=torch.int32, device=dev)
    items = torch.empty(n_dst + E // dcap + 1, 4, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        check(L.gg_csr_items(ptr(rowptr), n_dst, dcap, ptr(item_ptr), ptr(items), ptr(ws), ws_bytes, _stream()), 'gg_csr_items')
    if validate and int(status.item()) != 0:
        raise IndexError('edge_index holds an endpoint outside [0, N) (gg_csr_build: GG_ERANGE)')
    return EdgeCSR(rowptr, col, perm, n_src, n_dst, E, items, item_ptr)


def edge_wrap(csr, x_src, x_dst, out=None):
    """Per-edge periodic wrap codes (periodGATconv.py:209-210) in CSR order from the CURRENT positions (columns 0..2)."""
    if out is None:
        out = torch.empty(max(csr.n_edges, 1), dtype=torch.int32, device=x_dst.device)
    if csr.n_edges == 0:                       # an edge type that lost all its edges: nothing to classify
        return out
    with torch.cuda.device(x_dst.device):
        check(_lib.lib().gg_edge_wrap(ptr(x_src), x_src.stride(0), ptr(x_dst), x_dst.stride(0), ptr(csr.rowptr), ptr(csr.col),
                                      csr.n_dst, ptr(out), _stream()), 'gg_edge_wrap')
    return out


class CSRCache:
    """Small identity-keyed cache: hit iff the SAME tensor object is passed again, unmodified."""

    def __init__(self, capacity=16):
        self.capacity = capacity
        self._items = {}

    def get(self, edge_index, n_src, n_dst):
        key = id(edge_index)
        hit = self._items.get(key)
        if hit is not None:
            ref, version, shape, ns, nd, csr = hit
            if ref() is edge_index and version == edge_index._version and shape == tuple(edge_index.shape) \
                    and ns == n_src and nd == n_dst:
                return csr
        csr = build_csr(edge_index, n_src, n_dst)
        if len(self._items) >= self.capacity:
            for k in [k for k, v in self._items.items() if v[0]() is None] or list(self._items)[:1]:
                self._items.pop(k, None)
        self._items[key] = (weakref.ref(edge_index), edge_index._version, tuple(edge_index.shape), n_src, n_dst, csr)
        return csr


GLOBAL_CSR_CACHE = CSRCache()


def permute_to_csr(values, csr):
    """values[E] (original edge order) -> CSR order."""
    v = values.reshape(-1).contiguous()
    out = torch.empty_like(v)
    with torch.cuda.device(v.device):
        check(_lib.lib().gg_permute_f32(ptr(v), ptr(csr.perm), ptr(out), v.numel(), _stream()), 'gg_permute_f32')
    return out


def edge_length(x_src, x_dst, edge_index, csr=None):
    """Wrapped 2-D edge length (test.py:562-575). Returns ([E,1] original order, [E] CSR order or None)."""
    E = edge_index.shape[1]
    out = torch.empty(E, 1, dtype=torch.float32, device=x_src.device)
    out_csr = torch.empty(E, dtype=torch.float32, device=x_src.device) if csr is not None else None
    ei = edge_index.contiguous()
    with torch.cuda.device(x_src.device):
        check(_lib.lib().gg_edge_length(ptr(x_src), x_src.stride(0), ptr(x_dst), x_dst.stride(0), ptr(ei), E,
                                        ptr(csr.perm) if csr is not None else None, ptr(out), ptr(out_csr), _stream()),
              'gg_edge_length')
    return out, out_csr
