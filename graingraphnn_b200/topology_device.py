"""Row f1 on the device: the topology update of a rollout step without a host round trip (csrc/topology.cu,
csrc/topology_core.h — the routine the CPU suite checks against the reference's own `Cmodel.update` outputs).

`DeviceTopology` keeps the joint->joint and joint->grain edge arrays of the CALLER's numbering resident (int64 [2, cap], with head
room for the edges `delete_grain_index` appends) next to the engine, takes the event candidates straight from the engine's
`EventSelector` buffers and the predictions / joint rows from the engine's tensors, and leaves the engine with the new topology
(`set_topology`) and the new live-grain mask.  The host sees four integers per step (new edge counts, switches, eliminations).
No CPU fallback; nucleation (models.py:771-835) is not part of the device path (use `topology.topology_update` for it)."""
import ctypes

import torch

from . import _lib
from ._lib import check, ptr
from .engine import ET_GJ, ET_JG, ET_JJ

ERRORS = {1: 'position list overflow (a joint with more than 8 / a grain with more than 32 neighbours)',
          2: 'a grain to delete does not have exactly two joints (models.py:869 assert)',
          3: 'no grain shared across a side (models.py:673 KeyError / :925 unpacking)',
          4: 'a joint without three joint / grain neighbours', 5: 'sides and opposite grains of a vanishing grain do not match (models.py:681 assert)',
          6: 'a switch without exactly one growing grain per joint', 7: 'edge array capacity exceeded'}


class DeviceTopology:
    def __init__(self, engine, edge_index_dict, mask, headroom=None):
        """engine: RolloutEngine with the graph set and enable_event_selection() called; edge_index_dict / mask: the caller's."""
        self.eng = engine
        dev = engine.device
        self.dev = dev
        L = _lib.lib()
        cj, cg = ctypes.c_int32(), ctypes.c_int32()
        L.gg_topology_caps(ctypes.byref(cj), ctypes.byref(cg))
        self.cap_j, self.cap_g = cj.value, cg.value
        self.nj, self.ng = int(engine.xbuf['joint'].shape[0]), int(engine.xbuf['grain'].shape[0])
        pp, pq = edge_index_dict[ET_JJ].to(dev), edge_index_dict[ET_JG].to(dev)
        extra = headroom if headroom is not None else max(4096, pp.shape[1] // 16)
        self.pp = torch.full((2, pp.shape[1] + extra), -1, dtype=torch.int64, device=dev)
        self.pp[:, :pp.shape[1]] = pp
        self.pq = pq.clone().contiguous()
        self.n_pp, self.n_pq = int(pp.shape[1]), int(pq.shape[1])
        i32 = lambda *s: torch.zeros(*s, dtype=torch.int32, device=dev)   # noqa: E731
        self.lists = {'pp0': (i32(self.nj * self.cap_j), i32(self.nj)), 'pp1': (i32(self.nj * self.cap_j), i32(self.nj)),
                      'pq0': (i32(self.nj * self.cap_j), i32(self.nj)), 'pq1': (i32(self.ng * self.cap_g), i32(self.ng))}
        self.ahead_cnt = i32(self.nj)
        self.ahead_flag = torch.zeros(self.pp.shape[1], dtype=torch.uint8, device=dev)
        self.dirty_flag = torch.zeros(self.ng, dtype=torch.uint8, device=dev)
        self.dirty_list = i32(self.ng)
        self.act_g = torch.zeros(self.ng, dtype=torch.uint8, device=dev)
        self.act_j = torch.zeros(self.nj, dtype=torch.uint8, device=dev)
        self.status = i32(1)
        self.result = torch.zeros(8, dtype=torch.int64, device=dev)
        self.mask_g = mask['grain'].to(dev, torch.float32).reshape(-1).contiguous().clone()
        self.mask_j = mask['joint'].to(dev, torch.float32).reshape(-1).contiguous().clone()
        self.jrow = None if engine._node_rank is None else engine._node_rank['joint'].to(torch.int32).contiguous()
        self._bufs = None
        self.profile = False          # True: update() leaves per-phase device times in self.last_ms
        self.last_ms = None

    def edge_index(self):
        """The caller-numbered edge lists as the reference's cleanup leaves them (models.py:838-841)."""
        pp, pq = self.pp[:, :self.n_pp], self.pq[:, :self.n_pq]
        return {ET_JJ: pp, ET_JG: pq, ET_GJ: torch.flip(pq, dims=[0])}

    def _work(self, l1_cap, ge_cap):
        key = (l1_cap, ge_cap)
        if self._bufs is None or self._bufs[0] != key:
            L, dev = _lib.lib(), self.dev
            i32 = lambda n: torch.zeros(max(int(n), 1), dtype=torch.int32, device=dev)   # noqa: E731
            self._bufs = (key, {'scratch': i32(self.ng + 2 * (l1_cap + ge_cap) + 128), 'ge_sorted': i32(ge_cap), 'l1_work': i32(l1_cap),
                                'l1_logit': torch.zeros(max(l1_cap, 1), dtype=torch.float32, device=dev),
                                'switching': torch.zeros(max(l1_cap, 1), 2, dtype=torch.int64, device=dev),
                                'ge_out': i32(ge_cap + self.ng), 'work': i32(L.gg_topology_work_ints(l1_cap, ge_cap, self.ng))})
        return self._bufs[1]

    edge_prob_fn = staticmethod(torch.sigmoid)     # (a test may put the CPU's sigmoid here to compare with a host run bit for bit)

    @torch.no_grad()
    def update(self, pred):
        """Run the update for the step whose predictions are `pred` (the dict RolloutEngine.step returned: caller numbering).
        Returns {'switching_list' [S, 2] int64, 'grain_event' int64 ids (candidates + forced / swept), 'changed': bool}; the engine's
        topology, joint rows, prediction rows and masks are updated on the device."""
        eng, L, st = self.eng, _lib.lib(), torch.cuda.current_stream().cuda_stream
        sel = eng._events
        if sel is None:
            raise RuntimeError('enable_event_selection() first')
        (l1_count, l1_ids, l1_vals, l1_cap, _), (ge_count, ge_ids, ge_vals, ge_cap, _) = sel._buf['edge'], sel._buf['grain']
        w = self._work(l1_cap, ge_cap)
        # the switches run in the order of their PROBABILITIES (models.py:730-731 sorts sigmoid(edge_event)); logits that saturate or
        # collide in fp32 are ties there and keep their column order, so the kernel is handed torch's own sigmoid of the candidates
        l1_vals = self.edge_prob_fn(l1_vals)
        if self.pp.shape[1] - self.n_pp < 2 * (ge_cap + 2):                  # head room for the appended edges of this step
            grown = torch.full((2, self.n_pp + max(4096, 4 * (ge_cap + 2))), -1, dtype=torch.int64, device=self.dev)
            grown[:, :self.n_pp] = self.pp[:, :self.n_pp]
            self.pp = grown
            self.ahead_flag = torch.zeros(self.pp.shape[1], dtype=torch.uint8, device=self.dev)
        if eng.node_order is not None:                                       # the grain candidates are engine rows: the update works in the caller's ids
            ge_ids = eng.node_order['grain'].index_select(0, ge_ids.clamp(0, self.ng - 1).long()).to(torch.int32)
        yj = pred['joint'] if pred['joint'].is_contiguous() else pred['joint'].contiguous()
        yg = pred['grain']
        xj = eng.xbuf['joint']
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)] if self.profile else None
        if ev:
            ev[0].record()
        with torch.cuda.device(self.dev):
            for name, arr, n, n0, n1, c1 in (('pp', self.pp, self.n_pp, self.nj, self.nj, self.cap_j), ('pq', self.pq, self.n_pq, self.nj, self.ng, self.cap_g)):
                l0, c0 = self.lists[name + '0']
                l1_, c1_ = self.lists[name + '1']
                check(L.gg_topology_lists(ptr(arr), arr.shape[1], n, ptr(l0), ptr(c0), self.cap_j, n0, ptr(l1_), ptr(c1_), c1, n1, ptr(self.status), st),
                      'gg_topology_lists')
            if ev:
                ev[1].record()
            check(L.gg_topology_update(ptr(self.pp), self.pp.shape[1], self.n_pp, ptr(self.pq), self.pq.shape[1], self.n_pq,
                                       ptr(self.lists['pp0'][0]), ptr(self.lists['pp0'][1]), ptr(self.lists['pp1'][0]), ptr(self.lists['pp1'][1]),
                                       ptr(self.lists['pq0'][0]), ptr(self.lists['pq0'][1]), ptr(self.lists['pq1'][0]), ptr(self.lists['pq1'][1]),
                                       ptr(self.ahead_cnt), ptr(self.ahead_flag), ptr(xj), xj.stride(0), ptr(self.jrow), 6,
                                       ptr(yj), ptr(yg), yg.stride(0), ptr(self.mask_g), ptr(self.mask_j), ptr(self.act_g), ptr(self.act_j),
                                       self.nj, self.ng, ptr(ge_count), ptr(ge_ids), ptr(ge_vals), ge_cap, ptr(l1_count), ptr(l1_ids), ptr(l1_vals), l1_cap,
                                       ptr(self.dirty_flag), ptr(self.dirty_list), ptr(w['scratch']), ptr(w['ge_sorted']), ptr(w['l1_work']), ptr(w['l1_logit']),
                                       ptr(w['switching']), ptr(w['ge_out']), ptr(w['work']), ptr(self.result), st), 'gg_topology_update')
        if ev:
            ev[2].record()
        res = self.result.cpu().tolist()                                     # the only host read of the step: 7 integers
        n_pp, n_pq, n_sw, n_ge_out, err, n_ge_in, n_l1_in = res[:7]
        if int(self.status.item()):
            raise RuntimeError('gg_topology_lists: ' + ERRORS.get(int(self.status.item()), 'error'))
        if n_ge_in > ge_cap or n_l1_in > l1_cap:
            raise RuntimeError(f'{n_l1_in} edge / {n_ge_in} grain candidates exceed the selection buffers ({l1_cap} / {ge_cap}): '
                               f'enable_event_selection(cap=None) sizes them for the worst case')
        if err:
            raise RuntimeError('gg_topology_update: ' + ERRORS.get(err, f'error {err}'))
        if yj is not pred['joint']:
            pred['joint'].copy_(yj)
        changed = n_ge_out > 0 or n_sw > 0                                   # test.py:438 `topo`
        if changed:
            pp, pq = self.pp[:, :n_pp], self.pq[:, :n_pq]
            pp_new = pp[:, pp[0] != -1]                                      # cleanup (models.py:846-862): stable compaction
            pq_new = pq[:, pq[0] != -1]
            self.n_pp, self.n_pq = int(pp_new.shape[1]), int(pq_new.shape[1])
            self.pp[:, :self.n_pp] = pp_new
            self.pp[:, self.n_pp:n_pp] = -1
            self.pq = pq_new.contiguous()
            if ev:
                ev[3].record()
            eng.set_topology({ET_JJ: pp_new, ET_JG: self.pq, ET_GJ: torch.flip(self.pq, dims=[0]).contiguous()})
            eng.set_event_mask(self.mask_g)
        if ev and changed:
            ev[4].record()
            torch.cuda.synchronize()
            self.last_ms = {'lists': ev[0].elapsed_time(ev[1]), 'update_kernel': ev[1].elapsed_time(ev[2]), 'compaction': ev[2].elapsed_time(ev[3]),
                            'set_topology': ev[3].elapsed_time(ev[4])}
        return {'switching_list': w['switching'][:n_sw].clone(), 'grain_event': w['ge_out'][:n_ge_out].long(), 'changed': changed}
