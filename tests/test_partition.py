"""Slab partition.  CPU: the plan (ownership, halo lists, local numbering) and the halo exchange over a
world_size-2 gloo group (host plumbing; the pack step is injected because the product's pack kernel is CUDA-only).
GPU (-m gpu): a P-slab partition stepped in lockstep on one device equals the single-GPU engine."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import grain_oracle as orc
from util import ET, rel_err

from graingraphnn_b200.partition import HaloExchange, build_plan, local_features, owners_by_x, region_edges
from graingraphnn_b200.synth import honeycomb_graph, lattice_dims


def domain(px=4, py=2, seed=3):
    nx, ny = lattice_dims(px, py)
    x, ei, glob = honeycomb_graph(nx, ny, seed=seed, patches=(px, py), return_global=True)
    return x, ei, glob


def plans(x, ei, glob, world):
    owner = {t: owners_by_x(np.asarray(glob[t])[:, 0], world) for t in x}
    n = {t: int(v.shape[0]) for t, v in x.items()}
    return [build_plan(n, ei, owner, r, world) for r in range(world)], owner


@pytest.mark.parametrize('world', [2, 3, 8])
def test_plan_covers_every_node_and_edge_exactly_once(world):
    x, ei, glob = domain(8, 2)
    pl, owner = plans(x, ei, glob, world)
    for t in x:
        own = np.concatenate([p.own[t] for p in pl])
        assert np.array_equal(np.sort(own), np.arange(x[t].shape[0]))            # a partition of the nodes
        sizes = [p.n_own[t] for p in pl]
        assert max(sizes) - min(sizes) <= 0.05 * max(sizes)                       # balanced slabs
    for e in ET:
        gids = np.concatenate([p.edge_gid[e] for p in pl])
        assert np.array_equal(np.sort(gids), np.arange(ei[e].shape[1]))           # every edge on exactly one rank
        for p in pl:
            l2g = {t: np.concatenate([p.own[t], p.halo[t]]) for t in x}
            src, dst = p.edge_index[e]
            assert np.array_equal(l2g[e[0]][src], ei[e][0].numpy()[p.edge_gid[e]])
            assert np.array_equal(l2g[e[2]][dst], ei[e][1].numpy()[p.edge_gid[e]])
            assert (dst < p.n_own[e[2]]).all()                                    # targets are owned
            assert np.all(np.diff(p.edge_gid[e]) > 0)                             # original relative order kept
            if e[0] == e[2]:                                                      # global endpoints for the `src < dst` test
                assert np.array_equal(p.edge_global[e], ei[e].numpy()[:, p.edge_gid[e]])


@pytest.mark.parametrize('world', [2, 4])
def test_send_and_receive_lists_agree_between_ranks(world):
    x, ei, glob = domain(8, 2)
    pl, owner = plans(x, ei, glob, world)
    for r, p in enumerate(pl):
        for t in x:
            assert (owner[t][p.halo[t]] != r).all()
            for s, cnt in p.recv_cnt[t].items():
                off = p.recv_off[t][s]
                want = p.halo[t][off - p.n_own[t]: off - p.n_own[t] + cnt]        # global ids of the segment from rank s
                sent = pl[s].own[t][pl[s].send_idx[t][r]]                         # what rank s will pack for us
                assert np.array_equal(want, sent)
                assert pl[s].remote_off[t][r] == off                              # where a pushing peer must write
            assert p.n_local_max[t] == max(q.n_local[t] for q in pl)
    if world == 4:   # slabs only talk to their ring neighbours (periodic wrap links 0 and P-1)
        assert pl[0].peers == [1, 3] and pl[2].peers == [1, 3]


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _gloo_worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        x, ei, glob = domain(6, 2)
        owner = {t: owners_by_x(np.asarray(glob[t])[:, 0], world) for t in x}
        plan = build_plan({t: int(v.shape[0]) for t, v in x.items()}, ei, owner, rank, world)
        hx = HaloExchange(plan, 'cpu', 'nccl', pack=lambda src, idx, out=None: src.index_select(0, idx.long()))
        # a per-node tensor whose true value is a function of the GLOBAL id; halo rows start out as garbage
        items = []
        for width in (4, 96):
            it = {}
            for t in x:
                gid = torch.from_numpy(np.concatenate([plan.own[t], plan.halo[t]])).double()
                full = (gid[:, None] * 1e-3 + torch.arange(width)[None, :]).float()
                ten = full.clone()
                ten[plan.n_own[t]:] = -7.0
                it[t] = ten
            items.append(it)
        hx.exchange(items)
        ok = True
        for width, it in zip((4, 96), items):
            for t in x:
                gid = torch.from_numpy(np.concatenate([plan.own[t], plan.halo[t]])).double()
                full = (gid[:, None] * 1e-3 + torch.arange(width)[None, :]).float()
                ok &= bool(torch.equal(it[t], full))
        ret[rank] = ok and len(hx.bytes_sent_per_exchange) == 1 and hx.bytes_sent_per_exchange[0] > 0
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('world', [2, 3])
def test_halo_exchange_over_gloo(world):
    ctx = mp.get_context('spawn')
    ret = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=_gloo_worker, args=(r, world, port, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert all(ret.get(r) is True for r in range(world)), dict(ret)


def test_local_features_follow_local_numbering():
    x, ei, glob = domain(4, 2)
    pl, _ = plans(x, ei, glob, 2)
    xl = local_features(pl[1], x)
    for t in x:
        ids = np.concatenate([pl[1].own[t], pl[1].halo[t]])
        assert torch.equal(xl[t], x[t][torch.from_numpy(ids)])


@pytest.mark.parametrize('world', [2, 3])
def test_region_edges_give_every_owned_grain_its_joints_in_the_global_dict_order(world):
    """Geometry feedback on slabs (row f2): each rank holds, for the grains it owns, all their joints as local rows, keyed by
    the joint's first appearance in the UNDIVIDED grain->joint edge list (centres against the single-GPU engine: GPU test below)."""
    x, ei, glob = domain(8, 2)
    pl, owner = plans(x, ei, glob, world)
    gj = ei[ET[0]].numpy()
    first = np.full(x['joint'].shape[0], 2 ** 31 - 1, dtype=np.int64)
    np.minimum.at(first, gj[1], np.arange(gj.shape[1]))
    seen = 0
    for p in pl:
        edges, key = region_edges(p, ei, owner)
        l2g = {t: np.concatenate([p.own[t], p.halo[t]]) for t in x}
        assert (edges[0] < p.n_own['grain']).all() and (edges[1] < p.n_local['joint']).all()
        gg, jg = l2g['grain'][edges[0]], l2g['joint'][edges[1]]
        m = owner['grain'][gj[0]] == p.rank
        assert np.array_equal(gg, gj[0][m]) and np.array_equal(jg, gj[1][m])
        assert np.array_equal(key, first[jg])
        seen += edges.shape[1]
        # a grain's joints sorted by key come out in the order of the undivided graph's dict
        for g_loc in (0, p.n_own['grain'] // 2, p.n_own['grain'] - 1):
            sel = np.nonzero(edges[0] == g_loc)[0]
            mine = jg[sel][np.argsort(key[sel], kind='stable')]
            glob_sel = np.nonzero(gj[0] == p.own['grain'][g_loc])[0]
            assert np.array_equal(mine, gj[1][glob_sel][np.argsort(first[gj[1][glob_sel]], kind='stable')])
    assert seen == gj.shape[1]


def test_region_edges_reject_a_joint_that_is_not_local():
    x, ei, glob = domain(8, 2)
    pl, owner = plans(x, ei, glob, 2)
    bad = {e: v.clone() for e, v in ei.items()}
    far = int(np.argmin(np.abs(np.asarray(glob['joint'])[:, 0] - 0.75)))            # a joint deep inside slab 1 ...
    g0 = int(np.argmin(np.abs(np.asarray(glob['grain'])[:, 0] - 0.25)))              # ... hung on a grain deep inside slab 0
    assert owner['joint'][far] == 1 and owner['grain'][g0] == 0
    bad[ET[0]] = torch.cat([bad[ET[0]], torch.tensor([[g0], [far]])], 1)
    with pytest.raises(ValueError):
        region_edges(pl[0], bad, owner)



# ------------------------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
@pytest.mark.parametrize('world', [2, 4])
def test_partitioned_rollout_equals_single_gpu(world):
    """3 rollout steps: every owned output of every slab equals the single-GPU engine on the undivided graph (same kernels,
    same per-row edge order -> identical bits), and the slabs' union covers all nodes and jj edges."""
    from graingraphnn_b200.engine import RolloutEngine
    from graingraphnn_b200.partition import LocalSlabGroup
    dev = torch.device('cuda:0')
    x, ei, glob = domain(8, 2, seed=5)
    sd_r, sd_c = orc.synth_state_dict('regressor', 1), orc.synth_state_dict('classifier', 2)
    single = RolloutEngine.from_state_dicts(sd_r, sd_c, dev)
    single.set_graph({k: v.to(dev) for k, v in x.items()}, {k: v.to(dev) for k, v in ei.items()})
    group = LocalSlabGroup.build(sd_r, sd_c, x, ei, glob, world, dev)
    for _ in range(3):
        ref = single.step(6)
        group.step(6)
        got = {k: torch.full_like(ref[k], float('nan')) for k in ('joint', 'grain', 'grain_area', 'edge_event')}
        for e in group.engines:
            for k, (gid, val) in e.owned_predictions().items():
                got[k][torch.from_numpy(gid).to(dev)] = val
        for k in got:
            assert torch.isfinite(got[k]).all(), k
            assert torch.equal(got[k], ref[k]), (k, rel_err(got[k], ref[k]))
        for e in group.engines:      # resident features of the owned rows track the single-GPU state
            for t in ('grain', 'joint'):
                ids = torch.from_numpy(e.plan.own[t]).to(dev)
                assert torch.equal(e.x[t][:e.plan.n_own[t]], single.x[t][ids])


@pytest.mark.gpu
@pytest.mark.parametrize('world', [2, 3])
def test_partitioned_rollout_with_geometry_feedback_equals_single_gpu(world):
    """Row f2 on slabs: owned grains take the centres of their (owned + halo) joints, halo grains receive theirs through a
    fourth exchange; features, centres and predictions equal the single-GPU engine bit for bit over 3 steps."""
    from graingraphnn_b200.engine import RolloutEngine
    from graingraphnn_b200.partition import LocalSlabGroup
    dev = torch.device('cuda:0')
    x, ei, glob = domain(8, 2, seed=5)
    sd_r, sd_c = orc.synth_state_dict('regressor', 1), orc.synth_state_dict('classifier', 2)
    single = RolloutEngine.from_state_dicts(sd_r, sd_c, dev)
    single.set_graph({k: v.to(dev) for k, v in x.items()}, {k: v.to(dev) for k, v in ei.items()})
    single.enable_geometry_feedback()
    group = LocalSlabGroup.build(sd_r, sd_c, x, ei, glob, world, dev)
    for e in group.engines:
        e.enable_geometry_feedback()
    for _ in range(3):
        ref = single.step(6)
        group.step(6)
        for e in group.engines:
            pl = e.plan
            for k, (gid, val) in e.owned_predictions().items():
                assert torch.equal(val, ref[k][torch.from_numpy(gid).to(dev)]), k
            for t in ('grain', 'joint'):        # owned AND halo rows: the halo grains carry the centres their owners computed
                ids = torch.from_numpy(np.concatenate([pl.own[t], pl.halo[t]])).to(dev)
                assert torch.equal(e.x[t], single.x[t][ids]), t
            own_g = torch.from_numpy(pl.own['grain']).to(dev)
            assert torch.equal(torch.nan_to_num(e.centers, nan=-7.0), torch.nan_to_num(single.centers[own_g], nan=-7.0))
            for et in ET:
                gid = torch.from_numpy(pl.edge_gid[et]).to(dev)
                assert torch.equal(e.edge_attr[et], single.edge_attr[et][gid])


@pytest.mark.gpu
def test_partitioned_event_candidates_union_equals_single_gpu():
    """Row f1, first stage, on slabs: the union of the ranks' candidates (global ids) is the undivided graph's list."""
    from graingraphnn_b200.engine import RolloutEngine
    from graingraphnn_b200.partition import LocalSlabGroup
    dev = torch.device('cuda:0')
    x, ei, glob = domain(8, 2, seed=5)
    x['grain'][:, 3] *= 0.02                                  # some areas end up below the 1e-4 threshold
    mask = torch.ones(x['grain'].shape[0], 1)
    mask[::11] = 0
    sd_r, sd_c = orc.synth_state_dict('regressor', 1), orc.synth_state_dict('classifier', 2)
    single = RolloutEngine.from_state_dicts(sd_r, sd_c, dev)
    single.set_graph({k: v.to(dev) for k, v in x.items()}, {k: v.to(dev) for k, v in ei.items()})
    single.enable_event_selection(mask)
    group = LocalSlabGroup.build(sd_r, sd_c, x, ei, glob, 3, dev)
    for e in group.engines:
        e.enable_event_selection(mask)
    for _ in range(2):
        single.step(6)
        group.step(6)
        ref = single.fetch_events()
        got = [e.fetch_events() for e in group.engines]
        assert torch.equal(torch.sort(torch.cat([g['L1'] for g in got])).values, ref['L1'])
        ids = torch.cat([g['grain_event_ids'] for g in got])
        area = torch.cat([g['grain_event_area'] for g in got])
        order = torch.argsort(ids)
        assert torch.equal(ids[order], ref['grain_event_ids']) and torch.equal(area[order], ref['grain_event_area'])
        assert len(ref['L1']) > 0


@pytest.mark.gpu
@pytest.mark.parametrize('feedback', ['0', '1'])
@pytest.mark.parametrize('transport', ['nccl', 'p2p'])
def test_multi_gpu_partition(transport, feedback):
    """Real ranks, real halo exchange (needs >= 2 GPUs; skipped on the single-GPU box)."""
    import subprocess
    import sys
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip('needs >= 2 GPUs')
    n = 2 if n < 4 else 4
    env = dict(os.environ, GG_HALO=transport, GG_FEEDBACK=feedback)
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', f'--nproc-per-node={n}', '--master-addr', '127.0.0.1',
           '--master-port', str(_free_port()), os.path.join(os.path.dirname(os.path.abspath(__file__)), 'mgpu_check.py')]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and 'MGPU OK' in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
