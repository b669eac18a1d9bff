"""GPU: row f1 on the device — gg_topology_lists + gg_topology_update (csrc/topology.cu; the routine of csrc/topology_core.h that
tests/test_topology_core.py pins on the host) against the reference's OWN `Cmodel.update` outputs (tests/golden/
topology_golden.npz: edge lists position for position, moved joints bit for bit, masks, forced eliminations), and the rollout
driver with the device update against the driver with the host update."""
import ctypes
import os

import numpy as np
import pytest
import torch

import grain_oracle as orc
from test_topology_golden import CASES, _run, case
from util import ET, GOLDEN, load_graph

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def gold():
    return np.load(os.path.join(GOLDEN, 'topology_golden.npz'))


def device_update(x, ei, y, mask, active_grains, active_joints, threshold=0.6):
    """topology.topology_update's calling convention on the device entry points (candidates put into selection-style buffers)."""
    from graingraphnn_b200 import _lib
    from graingraphnn_b200._lib import check, ptr
    L, d = _lib.lib(), torch.device('cuda:0')
    st = torch.cuda.current_stream().cuda_stream
    nj, ng = x['joint'].shape[0], x['grain'].shape[0]
    prob = torch.sigmoid(y['edge_event'])
    pp0 = ei[ET[2]]
    L1 = ((prob > threshold) & (pp0[0] < pp0[1])).nonzero().view(-1)
    ge = y['grain_event']
    g = torch.Generator().manual_seed(0)
    shuffle = torch.randperm(L1.numel(), generator=g)                   # the selection kernel leaves its candidates in any order
    l1_ids = L1[shuffle].to(torch.int32).to(d)
    l1_vals = prob[L1[shuffle]].to(d)                                   # the candidates' probabilities: the order of the switches (models.py:730-731)
    gshuf = torch.randperm(ge.numel(), generator=g)
    ge_ids, ge_vals = ge[gshuf].to(torch.int32).to(d), y['grain_area'][ge[gshuf]].to(d)
    l1_cap, ge_cap = max(L1.numel(), 1), max(ge.numel(), 1)
    pad = lambda t, n, dt: torch.cat([t, torch.zeros(n - t.numel(), dtype=dt, device=d)]) if t.numel() < n else t   # noqa: E731
    l1_ids, l1_vals = pad(l1_ids, l1_cap, torch.int32), pad(l1_vals, l1_cap, torch.float32)
    ge_ids, ge_vals = pad(ge_ids, ge_cap, torch.int32), pad(ge_vals, ge_cap, torch.float32)
    l1_count = torch.tensor([L1.numel()], dtype=torch.int32, device=d)
    ge_count = torch.tensor([ge.numel()], dtype=torch.int32, device=d)
    cj, cg = ctypes.c_int32(), ctypes.c_int32()
    L.gg_topology_caps(ctypes.byref(cj), ctypes.byref(cg))
    cj, cg = cj.value, cg.value
    extra = 2 * (ge.numel() + 8) + ng // 4 + 64
    pp = torch.full((2, pp0.shape[1] + extra), -1, dtype=torch.int64, device=d)
    pp[:, :pp0.shape[1]] = pp0.to(d)
    pq = ei[ET[1]].to(d).contiguous()
    i32 = lambda n: torch.zeros(max(int(n), 1), dtype=torch.int32, device=d)   # noqa: E731
    lists = [(i32(nj * cj), i32(nj)), (i32(nj * cj), i32(nj)), (i32(nj * cj), i32(nj)), (i32(ng * cg), i32(ng))]
    status = i32(1)
    check(L.gg_topology_lists(ptr(pp), pp.shape[1], pp0.shape[1], ptr(lists[0][0]), ptr(lists[0][1]), cj, nj, ptr(lists[1][0]), ptr(lists[1][1]), cj, nj, ptr(status), st), 'gg_topology_lists')
    check(L.gg_topology_lists(ptr(pq), pq.shape[1], pq.shape[1], ptr(lists[2][0]), ptr(lists[2][1]), cj, nj, ptr(lists[3][0]), ptr(lists[3][1]), cg, ng, ptr(status), st), 'gg_topology_lists')
    assert int(status.item()) == 0
    xj, yj, yg = x['joint'].to(d).contiguous(), y['joint'].to(d).contiguous(), y['grain'].to(d).contiguous()
    mg, mj = mask['grain'].float().reshape(-1).to(d).contiguous(), mask['joint'].float().reshape(-1).to(d).contiguous()
    u8 = lambda n: torch.zeros(max(int(n), 1), dtype=torch.uint8, device=d)   # noqa: E731
    sw = torch.zeros(l1_cap, 2, dtype=torch.int64, device=d)
    ge_out = i32(ge_cap + ng)
    result = torch.zeros(8, dtype=torch.int64, device=d)
    # every work array is a named tensor: temporaries created inside the argument list would be freed (and their memory reused by the
    # next temporary) before the kernel runs
    ahead_cnt, ahead_flag, act_g, act_j = i32(nj), u8(pp.shape[1]), u8(ng), u8(nj)
    dirty_flag, dirty_list, scratch = u8(ng), i32(ng), i32(ng + 2 * (l1_cap + ge_cap) + 128)
    ge_sorted, l1_work, l1_logit = i32(ge_cap), i32(l1_cap), torch.zeros(l1_cap, dtype=torch.float32, device=d)
    work = i32(L.gg_topology_work_ints(l1_cap, ge_cap, ng))
    check(L.gg_topology_update(ptr(pp), pp.shape[1], pp0.shape[1], ptr(pq), pq.shape[1], pq.shape[1],
                               ptr(lists[0][0]), ptr(lists[0][1]), ptr(lists[1][0]), ptr(lists[1][1]), ptr(lists[2][0]), ptr(lists[2][1]), ptr(lists[3][0]), ptr(lists[3][1]),
                               ptr(ahead_cnt), ptr(ahead_flag), ptr(xj), xj.stride(0), None, 6, ptr(yj), ptr(yg), yg.stride(0),
                               ptr(mg), ptr(mj), ptr(act_g), ptr(act_j), nj, ng, ptr(ge_count), ptr(ge_ids), ptr(ge_vals), ge_cap,
                               ptr(l1_count), ptr(l1_ids), ptr(l1_vals), l1_cap, ptr(dirty_flag), ptr(dirty_list), ptr(scratch),
                               ptr(ge_sorted), ptr(l1_work), ptr(l1_logit), ptr(sw), ptr(ge_out), ptr(work), ptr(result), st), 'gg_topology_update')
    n_pp, n_pq, n_sw, n_ge_out, err = result.cpu().tolist()[:5]
    if err:
        raise RuntimeError(f'gg_topology_update error {err}')
    x['joint'].copy_(xj.cpu()); y['joint'].copy_(yj.cpu())
    mask['grain'].copy_(mg.cpu().view(-1, 1).to(mask['grain'].dtype)); mask['joint'].copy_(mj.cpu().view(-1, 1).to(mask['joint'].dtype))
    y['grain_event'] = ge_out[:n_ge_out].cpu().long()
    ppc, pqc = pp[:, :n_pp].cpu(), pq[:, :n_pq].cpu()
    out = {ET[2]: ppc[:, ppc[0] != -1], ET[1]: pqc[:, pqc[0] != -1]}
    out[ET[0]] = torch.flip(out[ET[1]], dims=[0])
    return x, out, sw[:n_sw].cpu()


@pytest.mark.parametrize('name,i', CASES)
def test_device_update_equals_the_reference_update(gold, name, i):
    c = case(gold, name, i)
    xo, eio, pairs, y, mask = _run(device_update, name, c)
    for et, short in ((ET[2], 'jj'), (ET[1], 'jg'), (ET[0], 'gj')):
        assert np.array_equal(eio[et].numpy(), c[f'ei_{short}_out']), short
    assert np.array_equal(pairs.numpy(), c['switching_list'])
    assert np.array_equal(y['grain_event'].numpy(), c['grain_event_out'])
    for t in ('joint', 'grain'):
        assert np.array_equal(xo[t].numpy(), c[f'x_{t}_out']), t
        assert np.array_equal(mask[t].numpy(), c[f'mask_{t}_out']), t
        assert np.array_equal(y[t].numpy(), c[f'y_{t}_out']), t


def test_device_update_reproduces_the_forced_eliminations(gold):
    k = 'c2_forced_'
    c = {f[len(k):]: gold[f] for f in gold.files if f.startswith(k)}
    x = {t: torch.from_numpy(c[f'x_{t}_in'].copy()) for t in ('joint', 'grain')}
    mask = {t: torch.from_numpy(c[f'mask_{t}_in'].copy()) for t in ('joint', 'grain')}
    ei = {et: torch.from_numpy(c[f'ei_{short}_in'].astype(np.int64)) for et, short in ((ET[0], 'gj'), (ET[1], 'jg'), (ET[2], 'jj'))}
    y = {'joint': torch.from_numpy(c['y_joint'].copy()), 'grain': torch.from_numpy(c['y_grain'].copy()),
         'edge_event': torch.from_numpy(c['y_edge_event']), 'grain_area': torch.from_numpy(c['y_grain_area'])}
    orc.regressor_update(x, y, span=0)
    _, y['grain_event'] = orc.event_candidates(y, ei[ET[2]], mask['grain'])
    _, eio, pairs = device_update(x, ei, y, mask, torch.from_numpy(c['active_grains']), torch.from_numpy(c['active_joints']))
    assert len(c['grain_event_out']) == len(c['grain_event_in']) + 2
    assert np.array_equal(y['grain_event'].numpy(), c['grain_event_out']) and np.array_equal(pairs.numpy(), c['switching_list'])
    for et, short in ((ET[2], 'jj'), (ET[1], 'jg'), (ET[0], 'gj')):
        assert np.array_equal(eio[et].numpy(), c[f'ei_{short}_out']), short
    for t in ('joint', 'grain'):
        assert np.array_equal(x[t].numpy(), c[f'x_{t}_out']) and np.array_equal(mask[t].numpy(), c[f'mask_{t}_out'])
        assert np.array_equal(y[t].numpy(), c[f'y_{t}_out'])


@pytest.mark.parametrize('rows', ['caller', 'morton'])
def test_rollout_driver_with_the_device_update_equals_the_host_update(rows):
    """Four frames with 3 eliminations and ~180 switches (the scenario of tests/test_rollout.py): the driver whose topology update
    runs on the device (gg_select_events buffers -> gg_topology_update -> set_topology, 7 integers to the host) against the driver
    with the host update."""
    from graingraphnn_b200.engine import RolloutEngine
    from graingraphnn_b200.rollout import RolloutDriver
    x, ei, ea = load_graph('c1')
    sd_r, sd_c = orc.synth_state_dict('regressor', 1), orc.synth_state_dict('classifier', 2)
    sd_c['lin2.bias'] = sd_c['lin2.bias'] - 0.613
    mask = {'grain': torch.ones(x['grain'].shape[0], 1), 'joint': torch.ones(x['joint'].shape[0], 1)}
    gp = {t: v[:, :2] for t, v in x.items()} if rows == 'morton' else None
    drv = {}
    for topo in ('host', 'device'):
        eng = RolloutEngine.from_state_dicts(sd_r, sd_c, torch.device('cuda:0'))
        drv[topo] = RolloutDriver(eng, x, ei, ea, mask, span=6, global_pos=gp, edge_threshold=0.6, area_threshold=0.0235, topology=topo)
    for s in range(4):
        for topo in ('host', 'device'):
            drv[topo].step()
        a, b = drv['host'].current_edge_index(), drv['device'].current_edge_index()
        for e in ET:
            assert torch.equal(a[e], b[e]), (s, e)
        for t in ('joint', 'grain'):
            xa = drv['host']._to_caller(t, drv['host'].eng.x[t]).cpu()
            xb = drv['device']._to_caller(t, drv['device'].eng.x[t]).cpu()
            assert torch.equal(xa, xb), (s, t)
    assert drv['host'].grain_event_list == drv['device'].grain_event_list and len(drv['host'].grain_event_list) == 3
    assert drv['host'].switch_count == drv['device'].switch_count > 0
    ma, mb = drv['host'].current_mask(), drv['device'].current_mask()
    assert torch.equal(ma['grain'].float(), mb['grain']) and torch.equal(ma['joint'].float(), mb['joint'])
    assert drv['device'].d2h_bytes < drv['host'].d2h_bytes / 20


@pytest.mark.parametrize('name,n_switch,n_vanish', [('c2', 150, 25), ('c2', 400, 60)])
def test_device_update_equals_the_host_update_on_dense_event_sets(name, n_switch, n_vanish):
    """Hundreds of events on the 1,043-grain graph (adjacent switching edges, grains of up to 8 sides, forced eliminations): long
    candidate lists, so the helper warp's sorts / list loops and the look-ahead warp run; every array equals the host update's
    (itself pinned by the reference's outputs), or both refuse the event set."""
    from test_topology_golden import _craft
    from graingraphnn_b200 import topology
    x0, ei, _ = load_graph(name)
    agreed = 0
    for seed in range(4):
        y0 = _craft(np.random.default_rng(7000 + seed), x0, ei, n_switch, n_vanish, 8)
        res = []
        for fn in (topology.topology_update, device_update):
            x = {k: v.clone() for k, v in x0.items()}
            y = {k: v.clone() for k, v in y0.items()}
            orc.regressor_update(x, y, span=0)
            _, y['grain_event'] = orc.event_candidates(y, ei[ET[2]])
            mask = {'grain': torch.ones(x['grain'].shape[0], 1), 'joint': torch.ones(x['joint'].shape[0], 1)}
            active = ((y['grain'][:, 0] > -10).nonzero().view(-1), (y['joint'][:, 0] > -10).nonzero().view(-1))
            try:
                _, eio, pairs = fn(x, ei, y, mask, *active)
                res.append((x, eio, pairs, y, mask))
            except (KeyError, AssertionError, ValueError, RuntimeError, IndexError):
                res.append(None)
        if res[0] is None or res[1] is None:
            assert res[0] is None and res[1] is None, seed
            continue
        (xa, ea_, pa, ya, ma), (xb, eb, pb, yb, mb) = res
        for et in ET:
            assert torch.equal(ea_[et], eb[et]), (seed, et)
        assert torch.equal(pa, pb) and torch.equal(ya['grain_event'], yb['grain_event'])
        for t in ('joint', 'grain'):
            assert torch.equal(xa[t], xb[t]) and torch.equal(ma[t], mb[t]) and torch.equal(ya[t], yb[t]), (seed, t)
        agreed += 1
    assert agreed >= 1
