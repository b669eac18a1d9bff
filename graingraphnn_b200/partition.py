"""Slab partition of a large grain domain across the GPUs of one box, with a halo exchange per message-passing hop
(SURVEY.md §8e).

Every PeriodConv is a 1-hop operator (periodGATconv.py:174: x_j = x_src[ei[0]]), and all gates of a cell read the same
cat([X, h]) (heteropgclstm.py:112), so a rank that owns a set of target nodes needs, besides its own rows, the rows of
the SOURCE endpoints of its in-edges that live on other ranks — the halo.  Ownership is by the node's GLOBAL x
coordinate ((x_patch + domain_offset) / domain_factor, test.py:43-44), never by the stored patch coordinate, which is
wrapped mod 1 (test.py:29-55); the periodic boundary makes slab P-1 a neighbour of slab 0 like any other.

Local numbering on a rank: [owned nodes, ascending global id | halo from rank 0 | halo from rank 1 | ...], so every
halo segment is a contiguous row range that the owner's rows are received (or pushed) straight into.  A rank holds
exactly the edges whose TARGET it owns, in their original relative order, so per-edge outputs map back to global edge
ids through `edge_gid`.

Three exchanges per rollout step (engine.RolloutEngine._step_gen): encoder h of both models -> decoder; decoder h of the
classifier's joints -> edge-event head (models.py:602 reads h[src]); updated X -> edge-length rebuild / next step.

Two transports:
  * 'nccl'  — pack kernel (gg_gather_rows) + grouped ncclSend/ncclRecv through torch.distributed (also runs over gloo for
              the CPU tests of the host logic);
  * 'p2p'   — the pack kernel stores straight into the peer's halo rows through NVLink peer pointers of a symmetric
              allocation (no staging, no NCCL on the data path), followed by one symmetric-memory barrier.
"""
import os

import numpy as np
import torch

from . import _lib
from ._lib import check, ptr
from .engine import RolloutEngine
from .graph import _stream


# ------------------------------------------------------------------------------------------------------ the plan
class SlabPlan:
    """Host-side (numpy) description of one rank's share of the global graph.  Deterministic: every rank derives the
    same global tables from the same inputs, then keeps only its own view."""

    def __init__(self, rank, world):
        self.rank, self.world = rank, world
        self.own, self.halo = {}, {}             # node type -> global ids
        self.n_own, self.n_local = {}, {}
        self.recv_off, self.recv_cnt = {}, {}    # node type -> {peer: first local row / rows} of the halo segment from peer
        self.send_idx = {}                       # node type -> {peer: local (owned) row ids the peer needs, in its halo order}
        self.remote_off = {}                     # node type -> {peer: first row of OUR segment in the peer's local numbering}
        self.edge_index, self.edge_gid = {}, {}  # edge type -> local [2,E_loc] int64 / global edge ids [E_loc]
        self.n_local_max = {}                    # node type -> max over ranks of n_local (symmetric allocations)

    @property
    def peers(self):
        p = set()
        for t in self.send_idx:
            p |= set(self.send_idx[t]) | set(self.recv_cnt[t])
        return sorted(p)


def owners_by_x(global_x, world):
    """Slab index of every node from its global x fraction in [0, 1)."""
    gx = np.asarray(global_x, dtype=np.float64)
    return np.minimum((gx * world).astype(np.int64), world - 1).astype(np.int32)


def morton_key(pos, bits=12):
    """Morton (Z-curve) key of pos[:, :2] in [0, 1)^2 (numpy int64) — the row order inside a slab, see RolloutEngine.set_graph."""
    q = np.clip((np.asarray(pos, dtype=np.float64)[:, :2] * (1 << bits)).astype(np.int64), 0, (1 << bits) - 1)
    code = np.zeros(q.shape[0], dtype=np.int64)
    for b in range(bits):
        code |= ((q[:, 0] >> b) & 1) << (2 * b)
        code |= ((q[:, 1] >> b) & 1) << (2 * b + 1)
    return code


def build_plan(n_nodes, edge_index, owner, rank, world, order_key=None):
    """n_nodes: {type: N}; edge_index: {(s, r, d): int64 [2, E] (numpy or torch)}; owner: {type: int32 [N]};
    order_key (optional): {type: int64 [N]} — the owned rows of a slab are laid out by ascending key (ties: global id) instead of
    ascending global id (a Morton key keeps the targets of a gather tile spatially compact)."""
    plan = SlabPlan(rank, world)
    ei = {e: (v.numpy() if isinstance(v, torch.Tensor) else np.asarray(v)) for e, v in edge_index.items()}
    types = list(n_nodes)
    # need[t] = unique (destination rank, source node) pairs with a foreign source, over every edge type whose source type is t
    need = {}
    for t in types:
        keys = []
        for e, idx in ei.items():
            if e[0] != t:
                continue
            r_dst = owner[e[2]][idx[1]].astype(np.int64)
            r_src = owner[t][idx[0]].astype(np.int64)
            m = r_dst != r_src
            keys.append(r_dst[m] * n_nodes[t] + idx[0][m])
        k = np.unique(np.concatenate(keys)) if keys else np.zeros(0, dtype=np.int64)
        need[t] = (k // max(n_nodes[t], 1), k % max(n_nodes[t], 1))     # (needing rank, global source id), sorted by both
    for t in types:
        own_count = np.bincount(owner[t], minlength=world)
        needer, gid = need[t]
        src_owner = owner[t][gid].astype(np.int64)
        plan.own[t] = np.nonzero(owner[t] == rank)[0]
        if order_key is not None:
            plan.own[t] = plan.own[t][np.argsort(order_key[t][plan.own[t]], kind='stable')]
        plan.n_own[t] = int(plan.own[t].shape[0])
        # our halo: rows we need, ordered by (owner rank, global id)
        mine = needer == rank
        h_gid, h_own = gid[mine], src_owner[mine]
        order = np.lexsort((h_gid, h_own))
        h_gid, h_own = h_gid[order], h_own[order]
        plan.halo[t] = h_gid
        plan.n_local[t] = plan.n_own[t] + int(h_gid.shape[0])
        plan.recv_off[t], plan.recv_cnt[t] = {}, {}
        for s in range(world):
            c = int((h_own == s).sum())
            if c:
                plan.recv_off[t][s] = plan.n_own[t] + int(np.searchsorted(h_own, s))
                plan.recv_cnt[t][s] = c
        # what the others need from us, each in the needing rank's halo order (ascending global id within our segment)
        g2l = np.full(n_nodes[t], -1, dtype=np.int64)
        g2l[plan.own[t]] = np.arange(plan.n_own[t])
        g2l[h_gid] = plan.n_own[t] + np.arange(h_gid.shape[0])
        plan._g2l = getattr(plan, '_g2l', {})
        plan._g2l[t] = g2l
        plan.send_idx[t], plan.remote_off[t] = {}, {}
        ours = src_owner == rank
        for s in range(world):
            sel = ours & (needer == s)
            if sel.any():
                plan.send_idx[t][s] = g2l[np.sort(gid[sel])].astype(np.int32)
                before = (needer == s) & (src_owner < rank)
                plan.remote_off[t][s] = int(own_count[s]) + int(before.sum())
        per_rank_halo = np.bincount(needer, minlength=world) if needer.size else np.zeros(world, dtype=np.int64)
        plan.n_local_max[t] = int((own_count + per_rank_halo).max())
    for e, idx in ei.items():
        m = owner[e[2]][idx[1]] == rank
        gids = np.nonzero(m)[0]
        src = plan._g2l[e[0]][idx[0][gids]]
        dst = plan._g2l[e[2]][idx[1][gids]]
        assert (src >= 0).all() and (dst >= 0).all() and (dst < plan.n_own[e[2]]).all()
        plan.edge_index[e] = np.stack([src, dst])
        plan.edge_gid[e] = gids
        if e[0] == e[2]:       # same-type edges: the GLOBAL endpoints decide `src < dst` (models.py:629), local numbering does not
            plan.edge_global = getattr(plan, 'edge_global', {})
            plan.edge_global[e] = np.stack([idx[0][gids], idx[1][gids]]).astype(np.int64)
    return plan


def region_edges(plan, edge_index_dict, owner):
    """Geometry feedback on a slab (engine.region_feedback): the grain->joint edges of the grains this rank owns, in local
    numbering, and for each the GLOBAL place of its joint in the reference's joint2vertex dict (first appearance as a target
    of the undivided grain->joint edge list, graph_trajectory.py:1062-1080) — so a grain's joints are walked in the same
    order on every partition.  The joints of an owned grain are local rows (own or halo) because the joint->grain edges
    into owned grains are this rank's; raises if the two edge types are not each other's reverse."""
    gj = edge_index_dict[('grain', 'push', 'joint')]
    gj = gj.numpy() if isinstance(gj, torch.Tensor) else np.asarray(gj)
    n_joint = owner['joint'].shape[0]
    first = np.full(n_joint, np.iinfo(np.int32).max, dtype=np.int64)
    uj, pos = np.unique(gj[1], return_index=True)                      # index of the first occurrence of every target
    first[uj] = pos
    m = owner['grain'][gj[0]] == plan.rank
    g = plan._g2l['grain'][gj[0][m]]
    j = plan._g2l['joint'][gj[1][m]]
    if (j < 0).any():
        raise ValueError('a joint of an owned grain is neither owned nor halo: the joint->grain edges are not the reverse of '
                         'the grain->joint edges')
    return np.stack([g, j]).astype(np.int64), first[gj[1][m]].astype(np.int32)


def local_features(plan, x_dict):
    """Rows of the global feature tensors in this rank's local numbering (owned, then halo)."""
    out = {}
    for t, x in x_dict.items():
        ids = torch.from_numpy(np.concatenate([plan.own[t], plan.halo[t]]))
        out[t] = x.index_select(0, ids)
    return out


# ------------------------------------------------------------------------------------------------------ transports
def p2p_available():
    """True when torch's symmetric memory (CUDA peer mappings over NVLink) can back the 'p2p' halo transport."""
    try:
        import torch.distributed._symmetric_memory as symm  # noqa: F401
        return torch.cuda.is_available() and torch.cuda.device_count() > 1
    except Exception:
        return False


def _pack_rows_cuda(src, idx, out=None):
    """out[i, :] = src[idx[i], :] with the library's pack kernel; `out` may be a raw (pointer, ld) pair on a PEER device."""
    n, w = int(idx.numel()), src.shape[1]
    if out is None:
        out = torch.empty(n, w, dtype=src.dtype, device=src.device)
    if isinstance(out, tuple):
        optr, ldo = out
    else:
        optr, ldo = out.data_ptr(), out.stride(0)
    with torch.cuda.device(src.device):
        check(_lib.lib().gg_gather_rows(ptr(src), src.stride(0), ptr(idx), n, w, optr, ldo, _stream()), 'gg_gather_rows')
    return out


class HaloExchange:
    """exchange(items): items = list of {node type: tensor [n_local, W]}; fills every tensor's halo rows from the owners.

    transport 'nccl': torch.distributed point-to-point (NCCL on GPUs; gloo in the CPU tests, which inject `pack`).
    transport 'p2p' : tensors must come from `alloc()` (symmetric memory); rows are stored directly into the peer."""

    def __init__(self, plan, device, transport='nccl', group=None, pack=None):
        self.plan, self.device, self.transport, self.group = plan, torch.device(device), transport, group
        self.pack = pack or _pack_rows_cuda
        self.send_idx = {t: {s: torch.from_numpy(v).to(self.device) for s, v in d.items()} for t, d in plan.send_idx.items()}
        self._handles = {}
        self.bytes_sent_per_exchange = []
        if transport == 'p2p':
            import torch.distributed._symmetric_memory as symm
            self._symm = symm
        elif transport != 'nccl':
            raise ValueError(f'unknown halo transport {transport!r}')

    # -- allocation: plain for nccl, symmetric (same size on every rank) for p2p
    def alloc(self, node_type, width):
        if self.transport == 'nccl':
            return torch.empty(self.plan.n_local[node_type], width, dtype=torch.float32, device=self.device)
        import torch.distributed as dist
        rows = self.plan.n_local_max[node_type]
        t = self._symm.empty((rows, width), dtype=torch.float32, device=self.device)
        hdl = self._symm.rendezvous(t, self.group if self.group is not None else dist.group.WORLD)
        view = t[:self.plan.n_local[node_type]]
        self._handles[view.data_ptr()] = hdl
        self._keep = getattr(self, '_keep', []) + [t]
        return view

    def exchange(self, items):
        if self.plan.world == 1:
            return
        if self.transport == 'p2p':
            return self._exchange_p2p(items)
        import torch.distributed as dist
        ops, keep, nbytes = [], [], 0
        for it in items:
            for t, ten in it.items():
                assert ten.is_contiguous()
                for s, idx in self.send_idx[t].items():
                    buf = self.pack(ten, idx)
                    keep.append(buf)
                    nbytes += buf.numel() * 4
                    ops.append(dist.P2POp(dist.isend, buf, s, self.group))
        for it in items:
            for t, ten in it.items():
                for s, cnt in self.plan.recv_cnt[t].items():
                    off = self.plan.recv_off[t][s]
                    ops.append(dist.P2POp(dist.irecv, ten[off:off + cnt], s, self.group))
        self.bytes_sent_per_exchange.append(nbytes)
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        del keep

    def _exchange_p2p(self, items):
        nbytes, hdl = 0, None
        for it in items:
            for t, ten in it.items():
                hdl = self._handles.get(ten.data_ptr())
                if hdl is None:
                    raise RuntimeError("p2p halo transport needs tensors from HaloExchange.alloc() (symmetric memory)")
                w = ten.shape[1]
                for s, idx in self.send_idx[t].items():
                    base = int(hdl.buffer_ptrs[s]) + self.plan.remote_off[t][s] * ten.stride(0) * 4
                    self.pack(ten, idx, (base, ten.stride(0)))          # stores travel over NVLink into the peer's halo rows
                    nbytes += idx.numel() * w * 4
        self.bytes_sent_per_exchange.append(nbytes)
        if hdl is not None:
            hdl.barrier(channel=0)                                       # every rank's rows have landed before anyone reads


# ------------------------------------------------------------------------------------------------------ the engine
class PartitionedEngine(RolloutEngine):
    """RolloutEngine on one slab of a global graph; `step()` runs the same kernels on the local rows and performs the
    three halo exchanges.  One process per GPU (torch.distributed)."""

    def __init__(self, regressor, classifier, device='cuda'):
        super().__init__(regressor, classifier, device)
        self.plan, self.halo = None, None

    def set_global_graph(self, x_dict, edge_index_dict, global_pos, rank, world, transport=None, group=None):
        """x_dict / edge_index_dict: the GLOBAL graph as CPU tensors (every rank passes the same data);
        global_pos[type][:, 0] = global x fraction in [0, 1) (slab coordinate)."""
        transport = transport or os.environ.get('GG_HALO', 'nccl')
        if transport == 'auto':
            transport = 'p2p' if p2p_available() else 'nccl'
        owner = {t: owners_by_x(np.asarray(global_pos[t])[:, 0], world) for t in x_dict}
        n_nodes = {t: int(v.shape[0]) for t, v in x_dict.items()}
        key = None if os.environ.get('GG_SLAB_ORDER', 'morton') != 'morton' else {t: morton_key(np.asarray(global_pos[t])) for t in x_dict}
        self.plan = build_plan(n_nodes, edge_index_dict, owner, rank, world, key)
        self.halo = HaloExchange(self.plan, self.device, transport, group)
        self.n_rows = dict(self.plan.n_own)           # kernels that WRITE per-node results stop at the owned rows
        self._region_src = (edge_index_dict, owner)   # region_edges() runs when the geometry feedback is first used
        self._region_edges = None
        self._events = self._event_edges = None       # bound to the previous plan's numbering: enable_event_selection() again
        xl = local_features(self.plan, x_dict)
        ei = {e: torch.from_numpy(v) for e, v in self.plan.edge_index.items()}
        self.set_graph({t: v.to(self.device) for t, v in xl.items()}, {e: v.to(self.device) for e, v in ei.items()})

    _region_edges = None
    _region_src = None

    def _region_index(self):
        from .geometry import RegionIndex
        if self._region_edges is None:
            self._region_edges = region_edges(self.plan, *self._region_src)
        edges, key = self._region_edges
        return RegionIndex(torch.from_numpy(edges).to(self.device), self.plan.n_own['grain'], self.plan.n_local['joint'],
                           edge_key=torch.from_numpy(key))

    def enable_geometry_feedback(self, joint_offset=None, domain_factor=1):
        """joint_offset: GLOBAL [Nj, 2] (as passed to the single-GPU engine); rows are taken in this rank's local numbering."""
        if joint_offset is not None:
            ids = torch.from_numpy(np.concatenate([self.plan.own['joint'], self.plan.halo['joint']]))
            joint_offset = joint_offset.cpu().index_select(0, ids)
        super().enable_geometry_feedback(joint_offset, domain_factor)

    def enable_event_selection(self, mask_grain=None, edge_threshold=0.6, area_threshold=1e-4, cap=None):
        """mask_grain: GLOBAL [Ng] or [Ng,1]; the owned rows are taken."""
        mask_grain = self._own_mask(mask_grain)
        super().enable_event_selection(mask_grain, edge_threshold, area_threshold, cap)
        self._event_edges = torch.from_numpy(self.plan.edge_global[('joint', 'connect', 'joint')]).to(self.device)

    def _own_mask(self, mask_grain):
        if mask_grain is None:
            return None
        m = mask_grain.cpu().reshape(mask_grain.shape[0], -1)[:, 0][torch.from_numpy(self.plan.own['grain'])]
        pad = self.plan.n_local['grain'] - m.shape[0]            # halo rows are never selected; the mask covers the local rows
        return torch.cat([m, m.new_zeros(pad)]) if pad > 0 else m

    def set_event_mask(self, mask_grain, local=False):
        """mask_grain: GLOBAL [Ng] or [Ng,1] (local=True: already this rank's rows)."""
        super().set_event_mask(mask_grain if local else self._own_mask(mask_grain))

    def fetch_events(self):
        """Candidates among the joint-joint edges and grains this rank owns, as GLOBAL ids (each edge / grain is owned by
        exactly one rank: the union over ranks, sorted the same way, is the undivided graph's list)."""
        ev = super().fetch_events()
        gid = torch.from_numpy(self.plan.edge_gid[('joint', 'connect', 'joint')])
        own = torch.from_numpy(self.plan.own['grain'])
        ev['L1'] = gid[ev['L1']]
        ev['grain_event'], ev['grain_event_ids'] = own[ev['grain_event']], own[ev['grain_event_ids']]
        return ev

    def alloc_rows(self, node_type, width):
        if self.halo is None:
            return super().alloc_rows(node_type, width)
        return self.halo.alloc(node_type, width)

    def alloc_rows_x(self, node_type, rows, width):
        if self.halo is None:
            return super().alloc_rows_x(node_type, rows, width)
        assert rows == self.plan.n_local[node_type]
        return self.halo.alloc(node_type, width)

    def _step_impl(self, span):
        """The step with its halo exchanges.  With the 'p2p' transport everything here is a kernel launch on the current stream
        (pack kernels that store into the peers' halo rows + the symmetric-memory barrier), so `capture()` records the whole
        partitioned step — exchanges included — into one CUDA graph per rank; the 'nccl' transport runs eagerly."""
        for items in self._step_gen(span):
            self.halo.exchange(items)
        return self.pred

    @torch.no_grad()
    def capture(self, span=6, warmup=2):
        if self.halo is not None and self.halo.transport != 'p2p':
            raise RuntimeError("capture() of a partitioned step needs the 'p2p' halo transport (GG_HALO=p2p): the NCCL path waits on the host")
        return super().capture(span, warmup)

    def counts(self):
        c = super().counts()
        c.update({'own_grain': self.plan.n_own['grain'], 'own_joint': self.plan.n_own['joint'],
                  'halo_grain': self.plan.n_local['grain'] - self.plan.n_own['grain'],
                  'halo_joint': self.plan.n_local['joint'] - self.plan.n_own['joint']})
        return c

    def owned_predictions(self):
        """Step outputs restricted to what this rank owns, with the global ids they belong to."""
        p, pl = self.pred, self.plan
        jj = ('joint', 'connect', 'joint')
        return {'joint': (pl.own['joint'], p['joint'][:pl.n_own['joint']]),
                'grain': (pl.own['grain'], p['grain'][:pl.n_own['grain']]),
                'grain_area': (pl.own['grain'], p['grain_area'][:pl.n_own['grain']]),
                'edge_event': (pl.edge_gid[jj], p['edge_event'])}


class LocalSlabGroup:
    """All slabs of a partition inside ONE process on ONE GPU, stepped in lockstep, halo rows copied with the pack kernel.
    Exists to check 'partitioned == single-GPU' without a multi-GPU box (tests) and to debug plans."""

    def __init__(self, engines):
        self.engines = engines

    @classmethod
    def build(cls, sd_r, sd_c, x_dict, edge_index_dict, global_pos, world, device='cuda'):
        engs = []
        for r in range(world):
            e = PartitionedEngine.from_state_dicts(sd_r, sd_c, device=device)
            owner = {t: owners_by_x(np.asarray(global_pos[t])[:, 0], world) for t in x_dict}
            e.plan = build_plan({t: int(v.shape[0]) for t, v in x_dict.items()}, edge_index_dict, owner, r, world)
            e.halo = None
            e.n_rows = dict(e.plan.n_own)
            e._region_src = (edge_index_dict, owner)
            xl = local_features(e.plan, x_dict)
            e.set_graph({t: v.to(e.device) for t, v in xl.items()},
                        {k: torch.from_numpy(v).to(e.device) for k, v in e.plan.edge_index.items()})
            engs.append(e)
        return cls(engs)

    @torch.no_grad()
    def step(self, span=6):
        gens = [e._step_gen(span) for e in self.engines]
        while True:
            reqs = []
            for g in gens:
                try:
                    reqs.append(next(g))
                except StopIteration:
                    reqs.append(None)
            if all(r is None for r in reqs):
                break
            assert all(r is not None for r in reqs), 'slabs fell out of lockstep'
            for r, eng in enumerate(self.engines):          # pull: rank r's halo rows <- owner s's rows
                for i, it in enumerate(reqs[r]):
                    for t, ten in it.items():
                        for s, cnt in eng.plan.recv_cnt[t].items():
                            off = eng.plan.recv_off[t][s]
                            src_eng = self.engines[s]
                            idx = torch.from_numpy(src_eng.plan.send_idx[t][r]).to(eng.device)
                            _pack_rows_cuda(reqs[s][i][t], idx, ten[off:off + cnt])
        return [e.pred for e in self.engines]
