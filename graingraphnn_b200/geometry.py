"""Geometry feedback of a rollout step on the device: grain centres from the joint positions.

The reference does this on the host after every NN step: `traj.GNN_update` (graph_trajectory.py:1010-1098) copies the
joint features to numpy, rebuilds joint2vertex from the grain->joint edges and calls `graph.update`
(graph_datastruct.py:672-708), whose per-grain Python loop unwraps the grain's joints across the periodic seam and
averages them; test.py:556-559 then assigns the centres to the grain features one grain at a time.  Here the index
(grain -> its joints in the reference's dict order) is built once per topology and one kernel per step writes the
centres straight into the resident grain rows.  Bit-exact against the reference's numpy arithmetic (tests/golden/
geometry_golden.npz is produced by the reference's own GNN_update).  No CPU fallback.
"""
import torch

from . import _lib
from ._lib import check, ptr
from .graph import build_csr


def _stream():
    return torch.cuda.current_stream().cuda_stream


class RegionIndex:
    """rowptr[n_grain+1], col[E] (joint ids), key[E] (place of the joint in the reference's joint2vertex dict) — int32.
    Built from the ('grain','push','joint') edge_index ([2,E] int64, CUDA); rebuilt whenever the topology changes."""

    __slots__ = ('rowptr', 'col', 'key', 'rank', 'col_sorted', 'n_grain', 'n_joint', 'n_edges')

    def __init__(self, gj_edge_index, n_grain, n_joint, edge_key=None):
        """edge_key (optional, int32 [E]): the dict position of each edge's joint, given by the caller — a slab of a
        partitioned domain passes the GLOBAL positions, so every rank walks a grain's joints in the same order as the
        undivided graph (partition.region_edges)."""
        if not gj_edge_index.is_cuda:
            raise RuntimeError('graingraphnn_b200 runs on CUDA tensors only (no CPU fallback)')
        ei = gj_edge_index.contiguous()
        E = ei.shape[1]
        dev = ei.device
        L = _lib.lib()
        by_grain = build_csr(torch.stack([ei[1], ei[0]]), n_joint, n_grain)    # rows = grains, col = joints, edge order kept
        self.rowptr, self.col = by_grain.rowptr, by_grain.col
        if edge_key is not None:
            self.rank = None
            self.key = edge_key.to(dev, torch.int32)[by_grain.perm.long()].contiguous() if E else torch.empty(1, dtype=torch.int32, device=dev)
        else:
            self.rank = torch.empty(max(n_joint, 1), dtype=torch.int32, device=dev)
            self.key = torch.empty(max(E, 1), dtype=torch.int32, device=dev)
            with torch.cuda.device(dev):
                check(L.gg_joint_rank(ptr(ei[1].contiguous()), E, n_joint, ptr(self.rank), _stream()), 'gg_joint_rank')
                check(L.gg_region_key(ptr(self.col), ptr(self.rank), E, ptr(self.key), _stream()), 'gg_region_key')
        # the joints of every grain in dict order: the per-step kernel walks them without keys
        self.col_sorted = torch.empty(max(E, 1), dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            check(L.gg_region_sort(ptr(self.rowptr), ptr(self.col), ptr(self.key), n_grain, ptr(self.col_sorted), _stream()), 'gg_region_sort')
        self.n_grain, self.n_joint, self.n_edges = n_grain, n_joint, E


def region_center(x_joint, index, x_grain=None, joint_offset=None, domain_factor=1, centers=None, want_centers=True, presorted=True):
    """Centres of all grains from the CURRENT joint rows (columns 0..1 of `x_joint`, fp32, even row stride).

    x_grain (optional): its columns 0..1 receive fp32(centre) — `(centre * domain_factor) % 1` on scaled patches
    (test.py:556-559); grains with <= 1 joint keep their coordinates (graph_datastruct.py:684).
    joint_offset [Nj,2] fp32 and domain_factor: global = (patch + offset) / factor (test.py:472-474).
    presorted=False walks the unsorted joint list by key (same results; kept as the cross-check of gg_region_sort).
    Returns float64 [n_grain, 2] centres (NaN rows for skipped grains), or None with want_centers=False."""
    if not x_joint.is_cuda:
        raise RuntimeError('graingraphnn_b200 runs on CUDA tensors only (no CPU fallback)')
    if x_joint.dtype != torch.float32 or x_joint.stride(1) != 1:
        raise ValueError('x_joint must be fp32 with unit column stride')
    if domain_factor > 1:
        if joint_offset is None:
            raise ValueError('domain_factor > 1 needs the per-joint patch offsets (test.py:43)')
        joint_offset = joint_offset.to(x_joint.device, torch.float32).contiguous()
        if tuple(joint_offset.shape) != (index.n_joint, 2):
            raise ValueError('joint_offset must be [n_joint, 2]')
    else:
        joint_offset = None
    if centers is None and want_centers:
        centers = torch.empty(index.n_grain, 2, dtype=torch.float64, device=x_joint.device)
    if x_grain is not None and (x_grain.dtype != torch.float32 or x_grain.stride(1) != 1 or x_grain.shape[0] < index.n_grain):
        raise ValueError('x_grain must be fp32 [>= n_grain, >= 2] with unit column stride')
    with torch.cuda.device(x_joint.device):
        check(_lib.lib().gg_region_center(ptr(x_joint), x_joint.stride(0), ptr(joint_offset), float(domain_factor),
                                          ptr(index.rowptr), ptr(index.col_sorted if presorted else index.col),
                                          None if presorted else ptr(index.key), index.n_grain,
                                          ptr(centers), ptr(x_grain), x_grain.stride(0) if x_grain is not None else 0,
                                          _stream()), 'gg_region_center')
    return centers


def area_bookkeeping(x_grain, mask_grain, gj_csr, index, lxd, patch_size=40.0, mesh_size=0.08, v_scale=20.0):
    """The QoI bookkeeping of `GNN_update` (graph_trajectory.py:1041-1051, :1100-1103) from the resident grain rows:
    -> (area_counts [Ng] float64: pixel-equivalent area of every live grain, normalised so the live grains tile the domain, NaN for
    dead ones; extraV [Ng] float64; vertex_area [Nj] float64: each grain's area shared equally among its joints, in um^2).
    x_grain [Ng, >= 5] fp32 (column 3 = area, 4 = extra volume); mask_grain [Ng] or [Ng, 1] (> 0: live) or None;
    gj_csr: EdgeCSR of ('grain','push','joint') (rows = joints); index: RegionIndex (rows = grains)."""
    if not x_grain.is_cuda:
        raise RuntimeError('graingraphnn_b200 runs on CUDA tensors only (no CPU fallback)')
    n_grain, n_joint = index.n_grain, index.n_joint
    dev = x_grain.device
    m = None
    if mask_grain is not None:
        m = mask_grain.to(dev, torch.float32).reshape(mask_grain.shape[0], -1)[:, 0].contiguous()
    s = patch_size / mesh_size + 1                                                  # :1043 (a float: no rounding here)
    area = x_grain[:n_grain, 3].double()
    area_sum = float(((area * m[:n_grain].double()) if m is not None else area).sum().item()) / (lxd / patch_size) ** 2
    counts = torch.empty(n_grain, dtype=torch.float64, device=dev)
    extra = torch.empty(n_grain, dtype=torch.float64, device=dev)
    varea = torch.empty(n_joint, dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        check(_lib.lib().gg_area_bookkeeping(ptr(x_grain), x_grain.stride(0), ptr(m), 1, n_grain, float(s), area_sum, float(v_scale),
                                             ptr(counts), ptr(extra), ptr(gj_csr.rowptr), ptr(gj_csr.col), ptr(index.rowptr), n_joint,
                                             float(mesh_size) ** 2, ptr(varea), _stream()), 'gg_area_bookkeeping')
    return counts, extra, varea
