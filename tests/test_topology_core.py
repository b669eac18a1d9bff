"""CPU: the device topology update — csrc/topology_core.h compiled for the host (tests/topology_host.cpp)
against the reference's OWN `Rmodel.update` + `Cmodel.update` outputs (tests/golden/topology_golden.npz) and against the
position-list host implementation (graingraphnn_b200/topology.py) on denser event sets.  The same header runs inside
gg_topology_update on the device (tests/test_gpu_topology.py)."""
import ctypes
import os
import shutil
import subprocess

import numpy as np
import pytest
import torch

import grain_oracle as orc
from test_topology_golden import CASES, _craft, _run, case
from util import ET, GOLDEN, load_graph

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def gold():
    return np.load(os.path.join(GOLDEN, 'topology_golden.npz'))


@pytest.fixture(scope='module')
def host_lib(tmp_path_factory):
    gxx = shutil.which('g++')
    if gxx is None:
        pytest.skip('g++ not available')
    out = str(tmp_path_factory.mktemp('topo') / 'libtopo_host.so')
    subprocess.run([gxx, '-O2', '-ffp-contract=off', '-shared', '-fPIC', '-o', out, os.path.join(ROOT, 'tests', 'topology_host.cpp')], check=True)
    return ctypes.CDLL(out)


def core_update(lib, preseed=1):
    """fn(x, ei, y, mask, active_grains, active_joints) -> (x, ei_out, pairs) with the calling convention of topology.topology_update,
    on the host build of the device routine."""
    def fn(x, ei, y, mask, active_grains, active_joints, threshold=0.6):
        nj, ng = x['joint'].shape[0], x['grain'].shape[0]
        logit = y['edge_event'].numpy()
        prob = torch.sigmoid(y['edge_event'])
        pp0 = ei[ET[2]]
        L1 = ((prob > threshold) & (pp0[0] < pp0[1])).nonzero().view(-1).numpy().astype(np.int32)      # models.py:627-629
        L1_logit = prob.numpy()[L1].astype(np.float32)                  # the order of the switches is by probability (models.py:730-731)
        ge = y['grain_event'].numpy().astype(np.int32)
        extra = 2 * (len(ge) + 8) + 2 * ng // 8 + 64                   # appended jj edges: two per deleted grain
        def grow(e, cap_extra):
            a = np.full((2, e.shape[1] + cap_extra), -1, dtype=np.int64)
            a[:, :e.shape[1]] = e.numpy()
            return a
        pp, pq = grow(ei[ET[2]], extra), grow(ei[ET[1]], 0)
        xj = np.ascontiguousarray(x['joint'].numpy()); yj = np.ascontiguousarray(y['joint'].numpy()); yg = np.ascontiguousarray(y['grain'].numpy())
        mg = mask['grain'].numpy().astype(np.float32).reshape(-1).copy(); mj = mask['joint'].numpy().astype(np.float32).reshape(-1).copy()
        ag = np.zeros(ng, np.uint8); ag[active_grains.numpy()] = 1
        aj = np.zeros(nj, np.uint8); aj[active_joints.numpy()] = 1
        sw = np.zeros((max(len(L1), 1), 2), np.int64)
        geo = np.zeros(len(ge) + ng + 8, np.int32)
        n_out = np.zeros(4, np.int64)
        P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
        L1c, L1l = L1.copy(), L1_logit.copy()
        rc = lib.topology_update_host(P(pp), ctypes.c_int64(pp.shape[1]), ctypes.c_int64(ei[ET[2]].shape[1]), P(pq), ctypes.c_int64(pq.shape[1]),
                                      ctypes.c_int64(pq.shape[1]), P(xj), xj.shape[1], 6, P(yj), P(yg), yg.shape[1], P(mg), P(mj), P(ag), P(aj), nj, ng,
                                      P(ge), len(ge), P(L1c), P(L1l), len(L1), P(sw), P(geo), P(n_out), preseed)
        if rc:
            raise RuntimeError(f'gg_topo_update error {rc}')
        x['joint'].copy_(torch.from_numpy(xj)); y['joint'].copy_(torch.from_numpy(yj))
        mask['grain'].copy_(torch.from_numpy(mg).view(-1, 1)); mask['joint'].copy_(torch.from_numpy(mj).view(-1, 1))
        y['grain_event'] = torch.from_numpy(geo[:n_out[3]].astype(np.int64))
        pp, pq = pp[:, :n_out[0]], pq[:, :n_out[1]]
        out = {ET[2]: torch.from_numpy(pp[:, pp[0] != -1].copy()), ET[1]: torch.from_numpy(pq[:, pq[0] != -1].copy())}   # cleanup, models.py:846-862
        out[ET[0]] = torch.flip(out[ET[1]], dims=[0])
        return x, out, torch.from_numpy(sw[:n_out[2]].copy())
    return fn


@pytest.mark.parametrize('preseed', [1, 0])
@pytest.mark.parametrize('name,i', CASES)
def test_device_routine_on_the_host_equals_the_reference_update(gold, host_lib, name, i, preseed):
    c = case(gold, name, i)
    xo, eio, pairs, y, mask = _run(core_update(host_lib, preseed), name, c)
    for et, short in ((ET[2], 'jj'), (ET[1], 'jg'), (ET[0], 'gj')):
        assert np.array_equal(eio[et].numpy(), c[f'ei_{short}_out']), short
    assert np.array_equal(pairs.numpy(), c['switching_list'])
    assert np.array_equal(y['grain_event'].numpy(), c['grain_event_out'])
    for t in ('joint', 'grain'):
        assert np.array_equal(xo[t].numpy(), c[f'x_{t}_out']), t
        assert np.array_equal(mask[t].numpy(), c[f'mask_{t}_out']), t
        assert np.array_equal(y[t].numpy(), c[f'y_{t}_out']), t


@pytest.mark.parametrize('name,n_switch,n_vanish,max_sides', [('c1', 25, 5, 7), ('c1', 60, 12, 8), ('c2', 150, 25, 8), ('c2', 400, 60, 8)])
def test_device_routine_on_the_host_equals_the_position_list_update(host_lib, name, n_switch, n_vanish, max_sides):
    from graingraphnn_b200 import topology
    rng = np.random.default_rng(n_switch * 31 + n_vanish)
    res = []
    for fn in (topology.topology_update, core_update(host_lib)):
        x, ei, _ = load_graph(name)
        y = _craft(np.random.default_rng(n_switch * 31 + n_vanish), x, ei, n_switch, n_vanish, max_sides)
        orc.regressor_update(x, y, span=0)
        _, y['grain_event'] = orc.event_candidates(y, ei[ET[2]])
        mask = {'grain': torch.ones(x['grain'].shape[0], 1), 'joint': torch.ones(x['joint'].shape[0], 1)}
        act_g = (y['grain'][:, 0] > -10).nonzero().view(-1)
        act_j = (y['joint'][:, 0] > -10).nonzero().view(-1)
        try:
            xo, eio, pairs = fn(x, ei, y, mask, act_g, act_j)
            res.append(('ok', xo, eio, pairs, y, mask))
        except Exception as exc:                                  # the reference's algorithm raises on some dense sets: both must
            res.append(('raised', type(exc).__name__))
    assert res[0][0] == res[1][0], (res[0][:2], res[1][:2])
    if res[0][0] == 'ok':
        (_, xa, ea, pa, ya, ma), (_, xb, eb, pb, yb, mb) = res
        for et in ET:
            assert torch.equal(ea[et], eb[et]), et
        assert torch.equal(pa, pb) and torch.equal(ya['grain_event'], yb['grain_event'])
        assert torch.equal(xa['joint'], xb['joint']) and torch.equal(ya['joint'], yb['joint'])
        assert torch.equal(ma['grain'], mb['grain']) and torch.equal(ma['joint'], mb['joint'])


def test_device_routine_on_the_host_reproduces_the_forced_eliminations(gold, host_lib):
    """The committed vector on which the reference's forced eliminations fire (fourth update of a chain on C2: 25 predicted + 2
    forced eliminations, 75 switches), masks carried over from the earlier updates."""
    k = 'c2_forced_'
    c = {f[len(k):]: gold[f] for f in gold.files if f.startswith(k)}
    x = {t: torch.from_numpy(c[f'x_{t}_in'].copy()) for t in ('joint', 'grain')}
    mask = {t: torch.from_numpy(c[f'mask_{t}_in'].copy()).float() for t in ('joint', 'grain')}
    ei = {et: torch.from_numpy(c[f'ei_{short}_in'].astype(np.int64)) for et, short in ((ET[0], 'gj'), (ET[1], 'jg'), (ET[2], 'jj'))}
    y = {'joint': torch.from_numpy(c['y_joint'].copy()), 'grain': torch.from_numpy(c['y_grain'].copy()),
         'edge_event': torch.from_numpy(c['y_edge_event']), 'grain_area': torch.from_numpy(c['y_grain_area'])}
    orc.regressor_update(x, y, span=0)
    _, y['grain_event'] = orc.event_candidates(y, ei[ET[2]], mask['grain'])
    _, eio, pairs = core_update(host_lib)(x, ei, y, mask, torch.from_numpy(c['active_grains']), torch.from_numpy(c['active_joints']))
    assert len(c['grain_event_out']) == len(c['grain_event_in']) + 2
    assert np.array_equal(y['grain_event'].numpy(), c['grain_event_out']) and np.array_equal(pairs.numpy(), c['switching_list'])
    for et, short in ((ET[2], 'jj'), (ET[1], 'jg'), (ET[0], 'gj')):
        assert np.array_equal(eio[et].numpy(), c[f'ei_{short}_out']), short
    for t in ('joint', 'grain'):
        assert np.array_equal(x[t].numpy(), c[f'x_{t}_out']) and np.array_equal(mask[t].numpy(), c[f'mask_{t}_out'].astype(np.float32))
        assert np.array_equal(y[t].numpy(), c[f'y_{t}_out'])


def test_device_routine_on_the_host_random_event_sets(host_lib):
    """40 random event sets on C1 / C2 (1..3 switches next to small grains, eliminations of 4- and 5-sided grains) against the
    position-list implementation: same arrays or the same refusal."""
    from graingraphnn_b200 import topology
    n_ok = 0
    for trial in range(40):
        name = 'c1' if trial % 2 == 0 else 'c2'
        n_switch, n_vanish = 1 + trial % 7, trial % 5
        res = []
        for fn in (topology.topology_update, core_update(host_lib)):
            x, ei, _ = load_graph(name)
            y = _craft(np.random.default_rng(1000 + trial), x, ei, n_switch, n_vanish, 5 + trial % 3)
            orc.regressor_update(x, y, span=0)
            _, y['grain_event'] = orc.event_candidates(y, ei[ET[2]])
            mask = {'grain': torch.ones(x['grain'].shape[0], 1), 'joint': torch.ones(x['joint'].shape[0], 1)}
            act_g = (y['grain'][:, 0] > -10).nonzero().view(-1)
            act_j = (y['joint'][:, 0] > -10).nonzero().view(-1)
            try:
                xo, eio, pairs = fn(x, ei, y, mask, act_g, act_j)
                res.append(('ok', xo, eio, pairs, y, mask))
            except Exception as exc:
                res.append(('raised', type(exc).__name__))
        assert res[0][0] == res[1][0], (trial, res[0][:2], res[1][:2])
        if res[0][0] == 'ok':
            n_ok += 1
            (_, xa, ea, pa, ya, ma), (_, xb, eb, pb, yb, mb) = res
            for et in ET:
                assert torch.equal(ea[et], eb[et]), (trial, et)
            assert torch.equal(pa, pb) and torch.equal(ya['grain_event'], yb['grain_event']), trial
            assert torch.equal(xa['joint'], xb['joint']) and torch.equal(ya['joint'], yb['joint']), trial
            assert torch.equal(ma['grain'], mb['grain']) and torch.equal(ma['joint'], mb['joint']), trial
    assert n_ok >= 30
