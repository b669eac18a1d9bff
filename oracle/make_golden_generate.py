"""Golden vectors of the reference's OWN generate mode (graph_trajectory.py:1289-1333) — run once in the build container,
where /root/reference exists:  python oracle/make_golden_generate.py
Writes tests/golden/generate_lxd{40,120,240}.npz: the HeteroGraph state the reference pickles (feature / edge-index /
edge-length / mask arrays), the area pixel counts and, for lxd 40, the rasterised alpha_field.  span = 6 is the value
the reference's nearest-neighbour lookup returns for G = 10, R = 2 (checked below against GR_train_grid.pkl).
TEST INFRASTRUCTURE ONLY."""
import contextlib
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shims  # noqa: E402

ET = [('grain', 'push', 'joint'), ('joint', 'pull', 'grain'), ('joint', 'connect', 'joint')]


def reference_generate(lxd, seed, G=10.0, R=2.0, span=None):
    ref_shims.install_plot_stubs()
    ref_shims.add_reference_to_path()
    cwd = os.getcwd()
    os.chdir('/root/reference')                      # GR_train_grid.pkl is opened by relative path
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            import dill
            import graph_trajectory as gt
            from scipy.interpolate import griddata
            traj = gt.graph_trajectory(lxd=lxd, seed=seed, frames=120, physical_params={'G': G, 'R': R})
            cur, counts = np.unique(traj.alpha_field, return_counts=True)
            traj.area_counts = dict(zip(cur, counts))
            traj.area_traj.append(traj.area_counts)
            traj.form_states_tensor(0)
            hg0 = traj.states[0]
            with open('GR_train_grid.pkl', 'rb') as inp:
                grid = dill.load(inp)
            G_ = (G - grid['G_min']) / (grid['G_max'] - grid['G_min'])
            R_ = (R - grid['R_min']) / (grid['R_max'] - grid['R_min'])
            hg0.span = griddata(np.array([grid['G'], grid['R']]).T, np.array(grid['span']), (G_, R_), method='nearest')
            if span is not None:
                assert int(hg0.span) == span, (hg0.span, span)
            hg0.form_gradient(prev=None, nxt=None, event_list=None, elim_list=None)
            hg0.append_history([])
    finally:
        os.chdir(cwd)
    return traj, hg0


def main():
    out = os.path.join(HERE, '..', 'tests', 'golden')
    for lxd, seed in ((40, 1), (120, 0), (240, 1)):
        traj, hg = reference_generate(lxd, seed, span=6)
        d = {'lxd': lxd, 'seed': seed, 'span': int(hg.span),
             'x_grain': hg.feature_dicts['grain'], 'x_joint': hg.feature_dicts['joint'],
             'mask_grain': hg.mask['grain'], 'mask_joint': hg.mask['joint'],
             'area_ids': np.array(list(traj.area_counts.keys()), dtype=np.int64),
             'area_counts': np.array(list(traj.area_counts.values()), dtype=np.int64)}
        for i, e in enumerate(ET):
            d[f'ei{i}'] = hg.edge_index_dicts[e].astype(np.int32)
            d[f'ew{i}'] = hg.edge_weight_dicts[e]
        if lxd == 40:
            d['alpha_field'] = traj.alpha_field.astype(np.int32)
        path = os.path.join(out, f'generate_lxd{lxd}.npz')
        np.savez_compressed(path, **d)
        print(path, os.path.getsize(path) // 1024, 'KiB', d['x_grain'].shape, d['x_joint'].shape, d['ei2'].shape)


if __name__ == '__main__':
    main()
