"""Generate tests/golden/geometry_golden.npz by running the REFERENCE's own geometry feedback —
graph_trajectory.GNN_update (graph_trajectory.py:1010-1103) -> graph.update (graph_datastruct.py:654-724) — imported
unmodified from /root/reference, followed by the grain-coordinate write-back of test.py:556-559 (restated: test.py is a
script).  Run once in the build container:  python oracle/make_golden_geometry.py

TEST INFRASTRUCTURE ONLY.  The trajectory object is created without its constructor (which
reads phase-field data) and given exactly the attributes GNN_update / update touch.
"""
import contextlib
import io
import os
import sys
from collections import defaultdict
from types import SimpleNamespace

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shims  # noqa: E402

ref_shims.install_plot_stubs()
ref_shims.add_reference_to_path()
from graph_trajectory import graph_trajectory  # noqa: E402  (reference)

OUT = os.path.join(HERE, '..', 'tests', 'golden')
GJ, JJ = ('grain', 'push', 'joint'), ('joint', 'connect', 'joint')


def bare_trajectory():
    t = object.__new__(graph_trajectory)
    t.BC = 'periodic'
    t.vertices = defaultdict(list)              # graph_datastruct.py:246-256
    t.vertex_neighbor = defaultdict(set)
    t.regions = defaultdict(list)
    t.region_coors = defaultdict(list)
    t.region_edge = defaultdict(set)
    t.region_center = defaultdict(list)
    t.quadruples = {}
    t.edges = []
    t.patch_size, t.mesh_size, t.lxd = 40, 0.08, 40
    t.states = [SimpleNamespace(targets_scaling={'grain': 1.0})]
    t.extraV_traj, t.area_traj = [], []
    return t


def reference_centers(x_joint_global, x_grain, gj, jj):
    """x_joint_global: fp32 [Nj, >=2] GLOBAL coordinates, as test.py:472-476 hands them to GNN_update."""
    t = bare_trajectory()
    n_g, n_j = x_grain.shape[0], x_joint_global.shape[0]
    mask = {'joint': torch.ones(n_j, 1), 'grain': torch.ones(n_g, 1)}
    ei = {GJ: torch.as_tensor(gj, dtype=torch.int64), JJ: torch.as_tensor(jj, dtype=torch.int64)}
    with contextlib.redirect_stdout(io.StringIO()):
        t.GNN_update(6, {'joint': x_joint_global, 'grain': x_grain}, mask, True, ei, False)
    c = np.full((n_g, 2), np.nan)
    for region, coor in t.region_center.items():
        c[region - 1] = coor
    return c


def writeback(x_grain, centers, factor):
    xg = x_grain.clone()
    for g in range(centers.shape[0]):          # test.py:556-559
        if not np.isnan(centers[g, 0]):
            xg[g, :2] = torch.FloatTensor(centers[g])
            if factor > 1:
                xg[g, :2] = (xg[g, :2] * factor) % 1
    return xg


def main():
    gold = {}
    rng = np.random.default_rng(7)
    for name, factor in (('c1', 1), ('c2', 3)):
        g = np.load(os.path.join(OUT, f'{name}_graph.npz'))
        xj = torch.from_numpy(g['x_joint'].copy())
        xg = torch.from_numpy(g['x_grain'].copy())
        # joints moved as a rollout step moves them (x_j[:, :2] += y_j / 5, |y_j| < 1 through tanh; models.py:510)
        xj[:, :2] += torch.from_numpy((rng.standard_normal((xj.shape[0], 2)) * 0.004).astype(np.float32))
        if factor > 1:
            # the fixture holds patch coordinates (test.py:29-44); the offsets come from the pickled global coordinates
            import dill
            with open('/root/reference/graphs/120_120/seed0_G10.0_R2.0_span6.pkl', 'rb') as f:
                raw = torch.FloatTensor(dill.load(f)[0].feature_dicts['joint'])[:, :2] * factor
            off = torch.floor(raw)                                                   # test.py:43
            assert torch.equal(raw - off, torch.from_numpy(g['x_joint'][:, :2]))
            glob = (xj[:, :2] + off) / factor                                        # test.py:474
        else:
            off = torch.zeros(xj.shape[0], 2)
            glob = xj[:, :2].clone()
        xin = xj.clone()
        xin[:, :2] = glob
        c = reference_centers(xin, xg, g['ei_gj'], g['ei_jj'])
        gold[f'{name}_x_joint'] = xj.numpy()
        gold[f'{name}_offset'] = off.numpy()
        gold[f'{name}_center'] = c
        gold[f'{name}_x_grain_out'] = writeback(xg, c, factor).numpy()
    # synthetic incidence: every joint touches 3 distinct grains, high grain degrees (the numpy mean switches to its
    # 8-lane pairwise form from 9 vertices on), edges in shuffled order, coordinates that straddle the periodic seam
    n_g, n_j = 48, 320
    tri, seen = [], set()
    while len(tri) < n_j:                     # distinct grain triples (two joints between the same three grains collapse
        t = tuple(sorted(rng.choice(44, 3, replace=False).tolist()))   # into one in the reference's joint2vertex dict)
        if t not in seen:
            seen.add(t)
            tri.append(t)
    tri = np.array(tri)
    tri[0] = [0, 1, 44]                       # grain 44: one vertex (skipped by graph_datastruct.py:684)
    tri[1] = [2, 3, 45]; tri[2] = [4, 5, 45]  # grain 45: two vertices; grains 46, 47: none
    gj = np.stack([tri.reshape(-1), np.repeat(np.arange(n_j), 3)])
    gj = gj[:, rng.permutation(gj.shape[1])]
    jj = np.stack([np.arange(n_j), (np.arange(n_j) + 1) % n_j])
    xj = torch.zeros(n_j, 8)
    xj[:, :2] = torch.from_numpy(((rng.random((n_j, 2)) * 0.3 + 0.85) % 1.0 - 0.002).astype(np.float32))
    xg = torch.from_numpy(rng.random((n_g, 11)).astype(np.float32))
    c = reference_centers(xj.clone(), xg, gj, jj)
    gold['syn_x_joint'], gold['syn_x_grain'] = xj.numpy(), xg.numpy()
    gold['syn_ei_gj'] = gj.astype(np.int32)
    gold['syn_center'] = c
    gold['syn_x_grain_out'] = writeback(xg, c, 1).numpy()
    np.savez_compressed(os.path.join(OUT, 'geometry_golden.npz'), **gold)
    print({k: v.shape for k, v in gold.items()})
    deg = np.bincount(gj[0], minlength=n_g)
    print('synthetic grain degrees', deg.min(), deg.max(), 'grains skipped:', int(np.isnan(c[:, 0]).sum()))


if __name__ == '__main__':
    main()
