"""The rollout loop of the reference's driver (test.py:353-577) on the resident engine — one object that strings together
what the reference does per frame:

  <1> regressor + classifier forward                     test.py:382-384      RolloutEngine.step (kernels b, c, c', d)
  <2> feature update, z advance                          :400-407             inside the step (gg_feature_update)
  <3> event candidates, topology update                  :418-426             gg_select_events on the device -> a few (id, value)
                                                                              pairs -> topology.topology_update on the host ->
                                                                              RolloutEngine.set_topology
  <4> geometry of the new tiling, QoIs                   :471-491, :523-533   gg_region_center (centres); event accounting;
                                                                              polygons + raster + layer error when a truth is given
  <5> grain coordinates <- centres, edge lengths         :556-575             gg_region_center write-back, gg_edge_refresh

Node features, indices and states stay on the GPU; what crosses PCIe per step is the candidate list, and — only on the steps
where something changes topology — the joint coordinates, the predictions the surgery reads and the new edge lists.
The QoIs the README quotes (`grain events hit rate a/b`, last-layer pointwise error, README.md:64-69) come out of
`RolloutDriver.qoi()`; they need the phase-field truth of the trajectory and, to mean anything, the shipped weights
(`weights.load_weights(regressor_pt, classifier_pt)`), both of which the caller supplies when they are at hand.
"""
import numpy as np
import torch

from .engine import ET_GJ, ET_JG, ET_JJ, RolloutEngine
from .generate import PATCH, Tiling, _update_init
from .topology import topology_update


def region_polygons(x_joint_global, gj_edge_index, mask_joint=None):
    """graph.update() as GNN_update calls it (graph_trajectory.py:1036-1098, graph_datastruct.py:654-724) -> (polygons, centres):
    polygons = {grain id (1-based): [n, 2] ccw vertex coordinates, unwrapped across the periodic seam} in the reference's
    `region_coors` dict order (the draw order of plot_polygons), centres [Ng, 2] (NaN: grains without a polygon).
    x_joint_global [Nj, 2]: joint coordinates in whole-domain units (test.py:472-474); gj_edge_index [2, E] grain -> joint."""
    xj = np.asarray(x_joint_global, dtype=np.float64)
    gj = np.asarray(gj_edge_index)
    nj = xj.shape[0]
    # vertex2joint[joint].add(grain + 1) in edge order; joint2vertex in order of each joint's first appearance (:1064-1085)
    first = np.full(nj, np.iinfo(np.int64).max, dtype=np.int64)
    first[gj[1][::-1]] = np.arange(gj.shape[1])[::-1]
    joints = np.flatnonzero(first < np.iinfo(np.int64).max)
    joints = joints[np.argsort(first[joints], kind='stable')]
    o = np.argsort(gj[1], kind='stable')
    cnt = np.bincount(gj[1], minlength=nj)
    if not (cnt[joints] == 3).all():
        raise AssertionError('a joint without exactly three grains (graph_trajectory.py:1068-1069)')
    start = np.zeros(nj + 1, dtype=np.int64)
    np.cumsum(cnt, out=start[1:])
    tri = np.sort(gj[0][o][start[joints][:, None] + np.arange(3)[None, :]] + 1, axis=1)
    t = Tiling()
    t.vertices, t.v2g, t.quadruples = xj, None, {}
    uniq, firstpos, inv = np.unique(tri, axis=0, return_index=True, return_inverse=True)
    if uniq.shape[0] != tri.shape[0]:                                  # repeated triple: first position, last vertex (dict semantics)
        inv = inv.reshape(-1)
        lastv = np.zeros(uniq.shape[0], dtype=np.int64)
        lastv[inv] = joints
        oo = np.argsort(firstpos, kind='stable')
        t.j2v_tri, t.j2v_vert = uniq[oo], lastv[oo]
    else:
        t.j2v_tri, t.j2v_vert = tri, joints
    _update_init(t)
    return t.polygons(), t.centers


class RolloutDriver:
    """engine: a RolloutEngine with weights loaded.  x / edge_index / edge_attr / mask: the driver's inputs after the loader and
    the patch scaling (generate.model_inputs, or the reference's own tensors), caller numbering.
    geometry: {'domain_factor', 'domain_offset' [Nj, 2] or 0} (test.py:310-312).  global_pos: see RolloutEngine.set_graph.
    truth (optional): {'grain_events': list over frames of sets of 1-based grain ids (traj.grain_events),
                       'alpha_pde': callable frame -> [s, s] int array (traj.alpha_pde_frames[:, :, frame].T), 'imagesize': s}
    raster: 'device' (gg_raster_polygons, row f4) or a callable (polygons, s) -> alpha_field [s, s] (graph_datastruct.py:553-610)."""

    def __init__(self, engine, x_dict, edge_index_dict, edge_attr_dict, mask, span=6, geometry=None, global_pos=None,
                 truth=None, raster='device', edge_threshold=0.6, area_threshold=1e-4, frames=121, ini_height=2.0, delta_z=0.4,
                 nucleation_density=0.0, lxd=None, topology='host'):
        self.eng, self.span, self.frames = engine, span, frames
        self.time_steps, self.step_seconds = False, []
        self.edge_threshold, self.area_threshold = edge_threshold, area_threshold
        self.geometry = geometry or {'domain_factor': 1, 'domain_offset': 0}
        self.factor = float(self.geometry.get('domain_factor', 1))
        off = self.geometry.get('domain_offset', 0)
        self.offset = off if isinstance(off, torch.Tensor) else None
        self.truth, self.raster = truth, raster
        self.ini_height, self.delta_z, self.nucleation_density = ini_height, delta_z, nucleation_density
        self.lxd = lxd if lxd is not None else PATCH * self.factor
        dev = engine.device
        self.edge_index = {e: v.cpu().clone() for e, v in edge_index_dict.items()}
        self.mask = {k: v.cpu().clone() for k, v in mask.items()}
        self.mask['joint'] = 1 + 0 * self.mask['joint']                                         # test.py:291
        engine.set_graph({k: v.to(dev) for k, v in x_dict.items()}, {k: v.to(dev) for k, v in edge_index_dict.items()},
                         None if edge_attr_dict is None else {k: v.to(dev) for k, v in edge_attr_dict.items()}, global_pos=global_pos)
        engine.enable_geometry_feedback(self.offset, self.factor, in_step=False)                # after the topology update, below
        engine.enable_event_selection(self.mask['grain'], edge_threshold, area_threshold)
        self.grain_event_list, self.grain_acc_list, self.layer_err_list = [], [(ini_height, 0, 0, 0)], []
        self.switch_count, self.topo_steps, self.d2h_bytes, self.h2d_bytes = 0, 0, 0, 0
        self.frame = 0
        self.topology = topology
        self._dtopo = None
        if topology == 'device':                                                                # row f1 without the host round trip
            if nucleation_density:
                raise NotImplementedError('nucleation is part of the host topology update only')
            from .topology_device import DeviceTopology
            self._dtopo = DeviceTopology(engine, self.edge_index, self.mask)
        elif topology != 'host':
            raise ValueError(topology)

    # ------------------------------------------------------------------------------------------------ order helpers
    def _to_caller(self, t, rows):
        r = self.eng._node_rank
        return rows if r is None else rows.index_select(0, r[t])

    def _to_engine(self, t, rows):
        o = self.eng.node_order
        return rows if o is None else rows.index_select(0, o[t])

    # ------------------------------------------------------------------------------------------------ one frame
    @torch.no_grad()
    def step(self):
        if getattr(self, 'time_steps', False):                                                  # (scripts/rollout.py --time-steps)
            import time
            torch.cuda.synchronize(self.eng.device)
            t0 = time.perf_counter()
            out = self._step()
            torch.cuda.synchronize(self.eng.device)
            self.step_seconds.append(time.perf_counter() - t0)
            return out
        return self._step()

    def _step(self):
        eng, span = self.eng, self.span
        self.frame += span
        frame = self.frame
        pred = eng.step(span)                                                                   # <1>, <2>
        if self._dtopo is not None:
            return self._step_device(pred, frame)
        ev = eng.fetch_events()                                                                 # <3> candidates only
        self.d2h_bytes += 16 + 8 * int(ev['L1'].numel() + ev['grain_event'].numel())
        grain_event, L1 = ev['grain_event'], ev['L1']
        pairs = torch.zeros(0, 2, dtype=torch.int64)
        if grain_event.numel() or L1.numel():
            # the surgery reads joint coordinates and predictions and moves joints: those arrays cross PCIe on such steps only
            x = {t: self._to_caller(t, eng.x[t]).cpu() for t in ('joint', 'grain')}
            y = {k: pred[k].cpu() for k in ('joint', 'grain', 'grain_area', 'edge_event')}
            y['grain_event'] = grain_event
            self.d2h_bytes += sum(v.numel() * 4 for v in x.values()) + sum(v.numel() * 4 for k, v in y.items() if k != 'grain_event')
            active_g = (y['grain'][:, 0] > -10).nonzero().view(-1)                               # models.py:502-503
            active_j = (y['joint'][:, 0] > -10).nonzero().view(-1)
            n_joint_live = float(self.mask['joint'].sum())
            nuc = self.nucleation_density * self.lxd * self.lxd * self.delta_z / max(n_joint_live, 1.0)   # test.py:424
            xj_before = x['joint'].clone()
            _, new_ei, pairs = topology_update(x, self.edge_index, y, self.mask, active_g, active_j, threshold=self.edge_threshold,
                                               L1=L1, nucleation_prob=float(nuc))
            grain_event = y['grain_event']                                                      # forced eliminations appended (models.py:757-759)
            if x['joint'].shape[0] != xj_before.shape[0]:
                raise NotImplementedError('nucleation grew the node set: rebuild the engine graph (set_graph) with the grown tensors')
            if grain_event.numel() or len(pairs):                                               # test.py:438 `topo`
                self.edge_index = new_ei
                dev = eng.device
                if not torch.equal(x['joint'], xj_before):                                      # joints moved by the switches (models.py:907, :992)
                    eng.x['joint'].copy_(self._to_engine('joint', x['joint'].to(dev)))
                    self.h2d_bytes += x['joint'].numel() * 4
                eng.set_topology({e: v.to(dev) for e, v in new_ei.items()})
                eng.set_event_mask(self.mask['grain'])
                self.h2d_bytes += sum(v.numel() * 8 for v in new_ei.values())
                self.topo_steps += 1
        self.grain_event_list.extend(int(g) for g in grain_event)                               # test.py:433
        self.switch_count += len(pairs)
        eng.region_feedback()                                                                   # <4> centres, <5> grain (x, y)
        eng.rebuild_edge_attr()                                                                 # <5> edge lengths (test.py:562-575)
        self._account(frame)
        return pred

    def _step_device(self, pred, frame):
        """<3> on the device (gg_topology_update): the host reads seven integers, plus the event ids for the accounting."""
        eng = self.eng
        out = self._dtopo.update(pred)
        self.d2h_bytes += 56 + 8 * int(out['grain_event'].numel())
        if out['changed']:
            self.topo_steps += 1
            self.edge_index = None                                                              # lives on the device: self._dtopo.edge_index()
        self.grain_event_list.extend(int(g) for g in out['grain_event'].cpu())
        self.switch_count += int(out['switching_list'].shape[0])
        eng.region_feedback()
        eng.rebuild_edge_attr()
        self._account(frame)
        return pred

    def _account(self, frame):
        height = self.ini_height + frame * self.delta_z
        if self.truth is not None:                                                              # test.py:480-491
            tr = self.truth['grain_events']
            ratio = self.truth.get('train_test_frame_ratio', 1)
            upto = frame // ratio + 1
            truth_set = set().union(*tr[:upto]) if len(tr) else set()
            truth_set = {i - 1 for i in truth_set}
            right = len(set(self.grain_event_list) & truth_set)
            self.grain_acc_list.append((height, len(truth_set), len(self.grain_event_list), right))
            if self.raster is not None and 'alpha_pde' in self.truth:                           # test.py:523-533
                self.layer_err_list.append((height, self.layer_error(frame // ratio)))

    def current_edge_index(self):
        """The caller-numbered edge lists of the current topology (CPU tensors)."""
        if self._dtopo is not None:
            return {e: v.cpu() for e, v in self._dtopo.edge_index().items()}
        return self.edge_index

    def current_mask(self):
        if self._dtopo is not None:
            return {'grain': self._dtopo.mask_g.cpu().view(-1, 1), 'joint': self._dtopo.mask_j.cpu().view(-1, 1)}
        return self.mask

    def run(self, frames=None):
        for _ in range(self.span, frames or self.frames, self.span):
            self.step()
        return self.qoi()

    # ------------------------------------------------------------------------------------------------ QoIs
    def polygons(self):
        """The grain polygons of the current tiling in whole-domain coordinates (what GNN_update + graph.update hand to plot_polygons)."""
        xj = self._to_caller('joint', self.eng.x['joint'])[:, :2].cpu()
        if self.factor > 1:
            xj = (xj + self.offset) / self.factor                                               # test.py:472-474
        return region_polygons(xj.numpy(), self.current_edge_index()[ET_GJ].numpy())[0]

    def layer_error(self, truth_frame):
        """plot_polygons + compute_error_layer (graph_datastruct.py:553-610, :346-348): on the device (raster.py) unless the
        caller supplied its own raster callable."""
        s = self.truth['imagesize']
        pde = self.truth['alpha_pde'](truth_frame)
        if self.raster == 'device':
            from . import raster
            alpha = raster.plot_polygons(self.polygons(), s, self.eng.device)
            self.alpha_field = alpha
            return raster.error_layer(torch.as_tensor(np.ascontiguousarray(pde)), alpha)
        alpha = self.raster(self.polygons(), s)
        self.alpha_field = alpha
        return float(np.sum(pde != alpha) / pde.size)

    def qoi(self):
        out = {'frames': self.frame, 'predicted_grain_events': len(self.grain_event_list), 'switches': self.switch_count,
               'topology_steps': self.topo_steps, 'd2h_bytes': self.d2h_bytes, 'h2d_bytes': self.h2d_bytes}
        if self.truth is not None and len(self.grain_acc_list) > 1:
            _, n_truth, n_pred, right = self.grain_acc_list[-1]
            out['grain_events_hit_rate'] = f'{right}/{n_truth}'                                 # README.md:68-69 "72/75", "644/704"
            out['grain_events_false_positive'] = n_pred - right
        if self.layer_err_list:
            out['last_layer_error'] = self.layer_err_list[-1][1]                                # README.md:68-69 "0.11", "0.18"
        return out
