"""RolloutEngine — the resident-on-GPU rollout step (the "nn-step" of SURVEY.md §8d).

One step = regressor forward + classifier forward (encoder + decoder cell each) + heads + in-place feature update +
edge-length rebuild, i.e. test.py:382-383, :400-407, :562-575 for a fixed topology.  Node features, CSR indices, edge
attributes, hidden states and all workspaces stay in HBM between steps; a step launches ~40 kernels and can be replayed
from a CUDA graph.  The host topology update of the reference (`Cmodel.update`, models.py:614-845) plugs in between steps
through `x`, `pred` and `set_topology()`.
"""
import torch

from . import _lib
from .cell import run_cell
from .graph import build_csr, edge_length, edge_wrap
from .heads import edge_head, feature_update, feature_update_batched, node_head
from .models import GrainNN_classifier, GrainNN_regressor
from .packing import pad4

import os
_EDGE_REFRESH = os.environ.get('GG_EDGE_REFRESH', '1') == '1'
_TWO_STREAMS = os.environ.get('GG_STREAMS', '2') == '2'

ET_GJ, ET_JG, ET_JJ = ('grain', 'push', 'joint'), ('joint', 'pull', 'grain'), ('joint', 'connect', 'joint')
DEFAULT_EDGE_TYPES = (ET_GJ, ET_JG, ET_JJ)


class _Hyper:
    def __init__(self, features, targets, layer_size, metadata, device):
        self.features, self.targets, self.layer_size, self.layers = features, targets, layer_size, 1
        self.metadata, self.device, self.out_win, self.window = metadata, device, 1, 1


def morton_order(pos, bits=12):
    """Row order along a Morton (Z) curve of pos[:, :2] in [0, 1)^2: int64 [N], order[i] = the node that takes row i."""
    q = (pos[:, :2].to(torch.float32) * (1 << bits)).long().clamp_(0, (1 << bits) - 1)
    code = torch.zeros(q.shape[0], dtype=torch.int64, device=pos.device)
    for b in range(bits):
        code |= ((q[:, 0] >> b) & 1) << (2 * b)
        code |= ((q[:, 1] >> b) & 1) << (2 * b + 1)
    return torch.sort(code, stable=True).indices


class RolloutEngine:
    def __init__(self, regressor, classifier, device='cuda'):
        self.device = torch.device(device)
        self.R, self.Cm = regressor.to(self.device).eval(), classifier.to(self.device).eval()
        self.C = regressor.out_channels
        self.edge_types = tuple(regressor.metadata[1])
        self.node_types = tuple(regressor.in_channels_dict)
        self.train_frames = 120
        self._graph = None
        self._work = {}
        self._work2 = {}          # second workspace set (GG_STREAMS=2: the classifier's cells run on their own stream)
        self._side = None
        self.x, self.xbuf, self.edge_index, self.edge_attr, self.ea_csr, self.csr, self.wrap = {}, {}, {}, {}, {}, {}, {}
        self.pred = {}
        self._scratch = torch.zeros(1, dtype=torch.int32, device=self.device)
        self._packs = None
        self.launches_per_step = 0

    # ---------------------------------------------------------------------------------------------------- setup
    @classmethod
    def from_state_dicts(cls, sd_regressor, sd_classifier, device='cuda', n_grain_feat=11, n_joint_feat=8,
                         edge_types=DEFAULT_EDGE_TYPES):
        C = sd_regressor['linear.grain.weight'].shape[1]
        hp = _Hyper({'grain': list(range(n_grain_feat)), 'joint': list(range(n_joint_feat))},
                    {'grain': [0, 1], 'joint': [0, 1]}, C, (['grain', 'joint'], list(edge_types)), 'cuda')
        R = GrainNN_regressor(hp)
        R.load_state_dict(sd_regressor)
        Cm = GrainNN_classifier(hp, R)
        Cm.load_state_dict(sd_classifier)
        return cls(R, Cm, device)

    def _pack(self):
        if self._packs is None:
            self._packs = {}
            for name, m in (('R', self.R), ('C', self.Cm)):
                enc, dec = m.gclstm_encoder.cell_list[0], m.gclstm_decoder.cell_list[0]
                self._packs[name] = (enc.packed(('i', 'c', 'o'), False, self.device, self.edge_types),
                                     dec.packed(('i', 'f', 'c', 'o'), True, self.device, self.edge_types))
        return self._packs

    node_order = None        # {node type: int64 [N]}: engine row -> caller's node id (None: rows are in the caller's order)
    _node_rank = None        # the inverse: caller's node id -> engine row

    def set_graph(self, x_dict, edge_index_dict, edge_attr_dict=None, global_pos=None):
        """Copy features into padded resident buffers, build the CSR of every edge type, take (or compute) edge lengths.

        global_pos (optional): {node type: [N, >= 2]} position of every node in the WHOLE domain, each coordinate in [0, 1)
        ((x_patch + domain_offset) / domain_factor, test.py:43-44, :474; the features themselves when the domain is one patch).
        When given, the engine keeps its rows along a Morton curve of these positions: the reference numbers grains in seed-lattice
        order and joints in first-seen order (graph_datastruct.py:118-160, :395-406), under which the targets of a gather tile share
        no source rows and every edge stages its own 1-2 KB row; along the curve a tile is a compact patch of the tiling and
        25-55 % of its edges find their source row already staged.  Everything the caller sees keeps the caller's numbering:
        `step()` returns node predictions in the caller's order, edge predictions in the original edge order, `set_topology`,
        `enable_geometry_feedback`, `enable_event_selection` / `fetch_events` take and return the caller's ids.  Only the resident
        rows `self.x[t]` / `host_features()` / `load_features()` are in engine order: row i is the caller's node `node_order[t][i]`."""
        self._graph = None
        self._state = None
        self._event_mask = None                       # belongs to the previous grain set (set_event_mask)
        self.node_order = self._node_rank = self._event_edges = None
        if global_pos is not None:
            self.node_order, self._node_rank = {}, {}
            for t in self.node_types:
                order = morton_order(global_pos[t].to(self.device))
                self.node_order[t] = order
                self._node_rank[t] = torch.empty_like(order).scatter_(0, order, torch.arange(order.shape[0], device=self.device))
            x_dict = {t: x_dict[t].to(self.device).index_select(0, self.node_order[t]) for t in self.node_types}
        for t in self.node_types:
            xt = x_dict[t].to(self.device, torch.float32)
            buf = self.alloc_rows_x(t, xt.shape[0], pad4(xt.shape[1]))
            buf.zero_()
            buf[:, :xt.shape[1]] = xt
            self.xbuf[t] = buf
            self.x[t] = buf[:, :xt.shape[1]]          # user-visible view; in-place edits land in the resident buffer
        self.set_topology(edge_index_dict, edge_attr_dict)

    def set_topology(self, edge_index_dict, edge_attr_dict=None):
        """edge_index_dict in the CALLER's node ids (and edge order: per-edge outputs follow it)."""
        self._graph = None
        self._region = None
        for e in self.edge_types:
            ei = edge_index_dict[e].to(self.device).contiguous()
            if self._node_rank is not None:           # engine rows of the end points; the edge order is the caller's
                if e == ET_JJ:
                    self._event_edges = ei            # `src < dst` of the event selection reads the caller's ids (models.py:629)
                ei = torch.stack([self._node_rank[e[0]][ei[0]], self._node_rank[e[2]][ei[1]]])
            self.edge_index[e] = ei
            self.csr[e] = build_csr(ei, self.xbuf[e[0]].shape[0], self.xbuf[e[2]].shape[0])
            self.edge_attr[e] = torch.empty(ei.shape[1], 1, dtype=torch.float32, device=self.device)
            self.ea_csr[e] = torch.empty(ei.shape[1], dtype=torch.float32, device=self.device)
            self.wrap[e] = torch.empty(max(ei.shape[1], 1), dtype=torch.int32, device=self.device)
        # The tile index of every (edge type, tile size) the cells will ask for is built HERE, on the current stream: run_cell
        # would otherwise build it lazily inside the first model's cell while the second model's cell, on the side stream,
        # finds it in the Python-side cache and launches its gather with nothing ordering it behind the build kernels.
        from .cell import tiled_ecap
        for encdec in self._pack().values():
            for pk in encdec:
                for e in self.edge_types:
                    ecap = tiled_ecap(pk, e)
                    if ecap:
                        self.csr[e].tiles(ecap)
        if self._event_mask is not None and self._event_mask.shape[0] < self.xbuf['grain'].shape[0]:
            self._event_mask = None                   # the grain set grew (nucleation): the old mask no longer covers it
        if edge_attr_dict is None:
            self.rebuild_edge_attr()
        else:
            L = _lib.lib()
            for e in self.edge_types:
                self.edge_attr[e].copy_(edge_attr_dict[e].to(self.device, torch.float32).reshape(-1, 1))
                _lib.check(L.gg_permute_f32(_lib.ptr(self.edge_attr[e]), _lib.ptr(self.csr[e].perm), _lib.ptr(self.ea_csr[e]),
                                            self.ea_csr[e].numel(), torch.cuda.current_stream().cuda_stream), 'gg_permute_f32')
            self.rebuild_edge_wrap()

    # ------------------------------------------------------------------------------------- event candidates (f1, first stage)
    _events = None
    _event_mask = None
    _event_edges = None      # [2, E] endpoints the `src < dst` test reads, when they differ from edge_index (slabs: global ids)

    def enable_event_selection(self, mask_grain=None, edge_threshold=0.6, area_threshold=1e-4, cap=None):
        """Every step also leaves the candidates of the host topology update on the device — the edges with
        sigmoid(edge_event) > edge_threshold and src < dst (models.py:627-629) and the live grains with predicted area below
        area_threshold (test.py:414) — so that `fetch_events()` copies a few (id, value) pairs instead of the full arrays.
        mask_grain: [Ng] or [Ng,1] fp32, > 0 for live grains (data['mask']['grain']); the reference reads the live mask every
        step (test.py:418), so call `set_event_mask()` after every topology update that changes it.
        cap: candidate slots per list; None = the worst case of the resident graph (E_jj / 2 edges with src < dst, every
        grain), so the buffers are never replaced under a captured step."""
        from .events import EventSelector
        if cap is None:
            ne = int(self.edge_index[ET_JJ].shape[1]) if ET_JJ in self.edge_index else 0
            ng = int(self.xbuf['grain'].shape[0]) if 'grain' in self.xbuf else 0
            cap_e, cap_g = max(ne // 2 + 1, 4096), max(ng, 4096)
        else:
            cap_e = cap_g = cap
        self._events = EventSelector(self.device, edge_threshold, area_threshold, cap_e, cap_g, on_grow=self._drop_graph)
        RolloutEngine.set_event_mask(self, mask_grain)   # (a subclass's override takes global rows)
        self._graph = None

    def _drop_graph(self):
        """A buffer the captured step writes was replaced: the graph holds stale pointers; later steps run eagerly until
        capture() is called again."""
        self._graph = None

    def set_event_mask(self, mask_grain):
        """The live-grain mask of the event selection (data['mask']['grain'], test.py:418): [N] or [N,1], > 0 = live.  It is
        copied, so call this again whenever the host topology update changes it (eliminations clear entries, nucleation
        appends grains).  A captured step is dropped: it reads the previous copy."""
        if mask_grain is None:
            self._event_mask = None
        else:
            m = mask_grain.to(self.device, torch.float32).reshape(mask_grain.shape[0], -1)[:, 0].contiguous()
            if self.node_order is not None:
                m = m.index_select(0, self.node_order['grain'])
            ng = int(self.xbuf['grain'].shape[0]) if 'grain' in self.xbuf else m.shape[0]
            if m.shape[0] < ng:
                raise ValueError(f'mask_grain has {m.shape[0]} rows for {ng} grains (refresh it after nucleation)')
            self._event_mask = m
        self._graph = None

    def fetch_events(self):
        """Host lists of the last step: {'L1' (ascending edge ids, models.py:629), 'L1_logit', 'grain_event' (sorted by area,
        test.py:416), ...}; ids are rows / edges of the graph this engine holds."""
        if self._events is None:
            raise RuntimeError('enable_event_selection() first')
        ev = self._events.fetch()
        if self.node_order is not None:               # grain rows -> the caller's grain ids
            order = self.node_order['grain'].cpu()
            ev['grain_event'], ev['grain_event_ids'] = order[ev['grain_event']], order[ev['grain_event_ids']]
        return ev

    # ------------------------------------------------------------------------------------- geometry feedback (f2)
    _geom = None
    _geom_in_step = True
    _region = None
    centers = None

    def enable_geometry_feedback(self, joint_offset=None, domain_factor=1, in_step=True):
        """From now on every step moves the grain coordinates to the centres of their joints before the edge lengths are
        rebuilt, as the reference's loop does on the host (traj.GNN_update -> graph.update, graph_datastruct.py:672-708, then
        test.py:556-559).  joint_offset [Nj,2] / domain_factor: the patch scaling of test.py:29-44 (global = (x + offset) / factor).
        self.centers holds the float64 centres of the last step (NaN rows: grains with <= 1 joint)."""
        if joint_offset is not None:
            joint_offset = joint_offset.to(self.device, torch.float32)
            if self.node_order is not None:
                joint_offset = joint_offset.index_select(0, self.node_order['joint'])
            joint_offset = joint_offset.contiguous()
        self._geom = (joint_offset, domain_factor)
        self._geom_in_step = bool(in_step)           # False: the caller runs region_feedback() itself, after its topology update (rollout.py)
        self._graph = None
        self._region = None

    def _region_index(self):
        from .geometry import RegionIndex
        return RegionIndex(self.edge_index[ET_GJ], self.xbuf['grain'].shape[0], self.xbuf['joint'].shape[0])

    def region_feedback(self):
        from .geometry import region_center
        if self._region is None:
            self._region = self._region_index()
            self.centers = torch.empty(self._region.n_grain, 2, dtype=torch.float64, device=self.device)
        off, factor = self._geom
        region_center(self.xbuf['joint'], self._region, self.xbuf['grain'], off, factor, self.centers)

    def rebuild_edge_wrap(self):
        """Per-edge periodic wrap codes (periodGATconv.py:209-210) of the CURRENT coordinates, shared by all 48 convs of a step."""
        for e in self.edge_types:
            edge_wrap(self.csr[e], self.xbuf[e[0]], self.xbuf[e[2]], self.wrap[e])

    def rebuild_edge_attr(self):
        """test.py:562-575 for every edge type, written in original and CSR order, together with the wrap codes of the new
        coordinates (periodGATconv.py:209-210).  One pass over the CSR rows per edge type (gg_edge_refresh); GG_EDGE_REFRESH=0:
        gg_edge_wrap + gg_edge_length (3 launches per edge type)."""
        L = _lib.lib()
        st = torch.cuda.current_stream().cuda_stream
        fused = _EDGE_REFRESH
        if not fused:
            self.rebuild_edge_wrap()
        for e in self.edge_types:
            xs, xd = self.xbuf[e[0]], self.xbuf[e[2]]
            g = self.csr[e]
            if fused:
                if g.n_edges:
                    _lib.check(L.gg_edge_refresh(_lib.ptr(xs), xs.stride(0), _lib.ptr(xd), xd.stride(0), _lib.ptr(g.rowptr), _lib.ptr(g.col),
                                                 _lib.ptr(g.perm), g.n_dst, _lib.ptr(self.wrap[e]), _lib.ptr(self.ea_csr[e]),
                                                 _lib.ptr(self.edge_attr[e]), st), 'gg_edge_refresh')
                continue
            ei = self.edge_index[e]
            _lib.check(L.gg_edge_length(_lib.ptr(xs), xs.stride(0), _lib.ptr(xd), xd.stride(0), _lib.ptr(ei), ei.shape[1],
                                        _lib.ptr(g.perm), _lib.ptr(self.edge_attr[e]), _lib.ptr(self.ea_csr[e]), st),
                       'gg_edge_length')

    # ---------------------------------------------------------------------------------------------------- step
    _state = None
    n_rows = None            # {node type: rows that per-node results are written for}; None = all (single-GPU engine)

    def alloc_rows_x(self, node_type, rows, width):
        return torch.empty(rows, width, dtype=torch.float32, device=self.device)

    def alloc_rows(self, node_type, width):
        """Storage of a per-node tensor other ranks read (h of the encoder, ...); the partitioned engine overrides it."""
        return torch.empty(self.xbuf[node_type].shape[0], width, dtype=torch.float32, device=self.device)

    def _states(self, name):
        st = self._state.get(name)
        if st is None:
            C = self.C
            st = {'he': {t: self.alloc_rows(t, C) for t in self.node_types}, 'ce': {},
                  'hd': {t: self.alloc_rows(t, C) for t in self.node_types}, 'cd': {}}
            self._state[name] = st
        return st

    def _step_gen(self, span):
        """One rollout step as a generator: yields, at each point where rows owned by other ranks are needed, the list of
        {node type: tensor} whose halo rows must be refreshed before execution continues.  The single-GPU engine just
        drains it; PartitionedEngine performs the exchanges (partition.py)."""
        if self._state is None:
            self._state = {}
        R, Cm = self.R, self.Cm
        packs, nr = self._pack(), self.n_rows
        sR, sC = self._states('R'), self._states('C')
        models = (('R', sR), ('C', sC))

        def both(fn):
            """The regressor's and the classifier's cells are independent until the heads.  They are issued on two
            streams (GG_STREAMS=1: one) (two branches of the captured graph): every kernel is one persistent CTA per SM, so nothing overlaps except
            the tail of one model's kernel with the head of the other's.  Each model then needs its own workspaces."""
            if not _TWO_STREAMS:
                for name, st in models:
                    fn(name, st, self._work)
                return
            main = torch.cuda.current_stream(self.device)
            if self._side is None:
                self._side = torch.cuda.Stream(device=self.device)
            self._side.wait_stream(main)
            fn('R', sR, self._work)
            with torch.cuda.stream(self._side):
                fn('C', sC, self._work2)
            main.wait_stream(self._side)

        # encoders of both models read only X (h0 = c0 = 0, models.py:237-238)
        both(lambda name, st, w: run_cell(packs[name][0], self.xbuf, None, None, self.csr, self.ea_csr, _lib.GG_GATE_LSTM0,
                                          st['he'], st['ce'], w, nr, self.wrap))
        yield [sR['he'], sC['he']]
        both(lambda name, st, w: run_cell(packs[name][1], self.xbuf, st['he'], st['ce'], self.csr, self.ea_csr, _lib.GG_GATE_LSTM,
                                          st['hd'], st['cd'], w, nr, self.wrap))
        yield [{'joint': sC['hd']['joint']}]                 # models.py:602 gathers h[src] of the joint-joint edges
        nj = None if nr is None else nr['joint']
        ng = None if nr is None else nr['grain']
        hd = sR['hd']
        yj, _ = node_head(hd['joint'], R.linear['joint'].weight, R.linear['joint'].bias, [1, 1], n_rows=nj)
        yg, area = node_head(hd['grain'], R.linear['grain'].weight, R.linear['grain'].bias, [1, 2],
                             area_in=self.xbuf['grain'][:, 3], area_scale=20.0, n_rows=ng)
        ev, ed = edge_head(sC['hd']['joint'], self.edge_index[ET_JJ], self.edge_attr[ET_JJ],
                           Cm.lin1.weight, Cm.lin1.bias, Cm.lin2.weight, Cm.lin2.bias)
        if self._events is not None:                         # row f1, first stage: only the event candidates leave the device
            self._events.select_edge_events(ev, self.edge_index[ET_JJ] if self._event_edges is None else self._event_edges)
            n_sel = area.shape[0] if ng is None else ng
            if self._event_mask is not None and self._event_mask.shape[0] < n_sel:
                raise RuntimeError(f'event mask covers {self._event_mask.shape[0]} of {n_sel} grains: call set_event_mask() after the topology update')
            self._events.select_grain_events(area[:n_sel], None if self._event_mask is None else self._event_mask[:n_sel])
        if isinstance(span, (tuple, list)):                  # ensemble: one span per graph of the block-diagonal batch
            dzj, dzg = self._dz_vectors(tuple(span))
            feature_update_batched(self.x['joint'], self.x['grain'], yj, yg, dzj, dzg,
                                   self.train_frames / (self.train_frames + 1), n_joint=nj, n_grain=ng)
        else:
            feature_update(self.x['joint'], self.x['grain'], yj, yg, span / (self.train_frames + 1),
                           self.train_frames / (self.train_frames + 1), self._scratch, n_joint=nj, n_grain=ng)
        yield [self.xbuf]                                    # moved coordinates of the halo -> edge lengths, next step
        if self._geom is not None and self._geom_in_step:    # row f2: grain centres follow their joints (test.py:471-476, :556-559)
            self.region_feedback()                           # owned grains, from owned + halo joints
            yield [{'grain': self.xbuf['grain']}]            # centres of the halo grains
        self.rebuild_edge_attr()
        if self.node_order is not None:                      # node predictions in the caller's numbering (3 small gathers)
            rj, rg = self._node_rank['joint'], self._node_rank['grain']
            self._pred_rows = {'joint': yj, 'grain': yg, 'grain_area': area}
            yj, yg, area = yj.index_select(0, rj), yg.index_select(0, rg), area.index_select(0, rg)
        self.pred = {'joint': yj, 'grain': yg, 'grain_area': area, 'edge_event': ev, 'edge': ed}

    graph_ptr = None         # {node type: [0, n_0, n_0 + n_1, ...]} of a block-diagonal batch (ensemble.EnsembleEngine)

    def _dz_vectors(self, spans):
        """Per-node z increment span_of_graph / (train_frames + 1) for a block-diagonal batch (cached per span tuple)."""
        if self.graph_ptr is None:
            raise ValueError('a per-graph span needs a batched graph (EnsembleEngine.set_graphs)')
        cache = self.__dict__.setdefault('_dz_cache', {})
        hit = cache.get(spans)
        if hit is None:
            out = []
            for t in ('joint', 'grain'):
                ptr_t = self.graph_ptr[t]
                if len(ptr_t) - 1 != len(spans):
                    raise ValueError(f'{len(spans)} spans for {len(ptr_t) - 1} graphs')
                counts = torch.tensor([ptr_t[i + 1] - ptr_t[i] for i in range(len(spans))])
                dz = torch.tensor([float(sp) / (self.train_frames + 1) for sp in spans], dtype=torch.float32)
                out.append(torch.repeat_interleave(dz, counts).to(self.device))
            hit = cache[spans] = tuple(out)
        return hit

    def _step_impl(self, span):
        for _ in self._step_gen(span):
            pass
        return self.pred

    @torch.no_grad()
    def step(self, span=6):
        """One rollout step on the resident graph. Returns the prediction dict (device tensors)."""
        if isinstance(span, list):
            span = tuple(span)
        if self._graph is not None and self._graph[0] == span:
            self._graph[1].replay()
            return self.pred
        with torch.cuda.device(self.device):
            out = self._step_impl(span)
        return out

    @torch.no_grad()
    def capture(self, span=6, warmup=2):
        """Capture the step into a CUDA graph (fixed topology): later step(span) calls replay it.
        NOTE: the warm-up runs advance the rollout state by max(warmup, 1) steps; the capture itself executes nothing."""
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            for _ in range(max(warmup, 1)):
                l0 = _lib.LAUNCHES[0]
                self._step_impl(span)
                self.launches_per_step = _lib.LAUNCHES[0] - l0
        torch.cuda.current_stream(self.device).wait_stream(side)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self._step_impl(span)
        self._graph = (span, g)
        return g

    # ---------------------------------------------------------------------------------------------------- host I/O
    def host_features(self):
        """CPU copies of the (unpadded) node features."""
        return {t: self.x[t].detach().cpu().contiguous() for t in self.node_types}

    def load_features(self, host_x):
        """H2D: overwrite the resident node features from (ideally pinned) host tensors, then refresh what derives from the
        coordinates (edge lengths, wrap codes), as test.py:556-575 does after the region-centre update."""
        for t in self.node_types:
            self.x[t].copy_(host_x[t], non_blocking=True)
        with torch.cuda.device(self.device):
            self.rebuild_edge_attr()

    def fetch_predictions(self, pred, out=None, keys=('joint', 'grain', 'grain_area', 'edge_event')):
        """D2H of the step outputs the host topology update consumes (models.py:626-628, test.py:418) into pinned buffers."""
        if out is None:
            out = {k: torch.empty(pred[k].shape, dtype=pred[k].dtype, pin_memory=True) for k in keys}
        for k in keys:
            out[k].copy_(pred[k], non_blocking=True)
        return out

    # ---------------------------------------------------------------------------------------------------- accounting
    def algorithmic_work(self):
        """Algorithmic bytes (gather) and flops (GEMMs) of ONE step on this engine's graph — DESIGN.md §4.
        gather, classic (decoder): per edge type and cell  4*G*C*(2 N_src + 2 N_dst) + 4*4*G*N_dst + 12 N_dst + 8 (N_dst+1) + 12 E
          (K, V once per source; Q, QX once and agg written once per target; target position; rowptr-equivalent item
          offsets; col + edge length + wrap code per edge)
        gather, raw-score: 4*(G*C + R)*N_src + 4*(G*C + R G)*N_dst + 12 N_dst + 8 (N_dst+1) + 12 E, R = 16 (encoder) or 32 + C (decoder)
          (raw input + V per source; Q' (R per gate) and agg per target)
        node_proj: 2 * N_t * K_t * ncols_t flop per node type and cell (K = F (+C with hidden state), unpadded)
        gate_update: 2 * N_t * G*C * (C * n_in_types + K_t) flop per node type and cell"""
        n = {t: self.xbuf[t].shape[0] for t in self.node_types}
        F = {t: self.x[t].shape[1] for t in self.node_types}
        C = self.C
        out = {'gg_pgat_gather': 0.0, 'gg_node_proj': 0.0, 'gg_gate_update': 0.0}
        for name in ('R', 'C'):
            for pk, with_h in zip(self._pack()[name], (False, True)):
                G = pk.G
                for e in self.edge_types:
                    ns, nd, E = n[e[0]], n[e[2]], int(self.edge_index[e].shape[1])
                    common = 12.0 * nd + 8.0 * (nd + 1) + 12.0 * E
                    if pk.raw_k:                  # the target's position rides inside Q' (no separate 12 bytes per target)
                        out['gg_pgat_gather'] += 4.0 * (G * C + pk.raw_k) * ns + 4.0 * (G * C + pk.raw_k * G) * nd + common - 12.0 * nd
                    else:
                        out['gg_pgat_gather'] += 4.0 * G * C * (2 * ns + 2 * nd) + 16.0 * G * nd + common
                for t in self.node_types:
                    K = F[t] + (C if with_h else 0)
                    out['gg_node_proj'] += 2.0 * n[t] * K * pk.ncols[t]
                    out['gg_gate_update'] += 2.0 * n[t] * G * C * (C * len(pk.into[t]) + K)
        return out

    def counts(self):
        ng, nj = self.xbuf['grain'].shape[0], self.xbuf['joint'].shape[0]
        return {'n_grain': ng, 'n_joint': nj, 'edges': sum(int(self.edge_index[e].shape[1]) for e in self.edge_types)}
