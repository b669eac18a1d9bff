#!/usr/bin/env python
"""bench.py — GrainGNN rollout throughput (edges/s, steps/s) on B200, with roofline and CPU baseline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--lxd L]

Workload: the periodic hex-Voronoi grain domain `graph_trajectory.py --mode=generate` builds (graingraphnn_b200/generate.py
— equal to the reference's output array for array at the sizes the reference reaches, tests/test_generate.py), G = 10,
R = 2, seed 1, through the loader + patch scaling of the rollout driver (test.py:29-55).
  N = 1: lxd = 1320 um (33 x 33 patches of 40 um, ~1.26 10^5 grains) — BASELINE config 3, single-B200 rollout;
  N = 2 / 4 / 8: lxd = 1880 / 2640 / 3720 um (the same ~1.25 10^5 grains per GPU; N = 8 is BASELINE config 4, 93 x 93
  patches, ~10^6 grains), x-slab partition with a halo exchange per message-passing hop (weak scaling);
  strong scaling: the 10^6-grain domain of config 4 on the N GPUs of this run (`strong_scaling`, GG_BENCH_STRONG=0 skips it).
One step = regressor + classifier forward (encoder + decoder cells) + heads + feature update + edge-length rebuild on a
fixed topology (the "nn-step" of SURVEY.md §8d).  Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

WEAK_LXD = {1: 1320, 2: 1880, 4: 2640, 8: 3720}
STRONG_LXD = 3720
CPU_SAMPLE_LXD = 240
CPU_BASELINE_LXD = 640           # `cpu_baseline` of the N = 1 line: 29,6xx grains, ~10 s of CPU work on 24 cores
SPAN = 6
ET = [('grain', 'push', 'joint'), ('joint', 'pull', 'grain'), ('joint', 'connect', 'joint')]
GEOM = {}          # lxd -> (per-joint patch offsets [Nj, 2], domain_factor) of the domains made so far (test.py:310-312)


def weak_lxd(n):
    return WEAK_LXD.get(n, int(round(1320 * n ** 0.5 / 40)) * 40)


def peaks():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            d = json.load(f)
        return {'hbm_gbs': d['hbm_gbs'], 'bf16_tflops': d['bf16_tflops'], 'bf16_sustained': d.get('bf16_tflops_sustained', d['bf16_tflops']),
                'source': 'measured'}
    except Exception:
        return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'bf16_sustained': 1400.0, 'source': 'fallback'}


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled every 100 ms while the timed region runs."""
    Q = 'index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,' \
        'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits', '-lms', '100',
                                          '-i', str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
            time.sleep(0.3)                       # the first sample must not miss a short timed region
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace('.', '').isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace('.', '').isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), r[5:9]):
                    if v.lower().startswith('active'):
                        reasons.add(name)
        return {'sm_mhz': statistics.median(sm) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


# ------------------------------------------------------------------------------------------------ workload
def make_domain(lxd, seed=1, rank=0, world=1):
    """(x, ei, ea, global positions, description) of the generate-mode domain, as the rollout driver hands it to the models.
    Cached under $GG_BENCH_CACHE (default /tmp/gg_bench_cache): the N = 1, 2, 4, 8 runs of one box share the 10^6-grain domain;
    with several ranks, rank 0 generates and the others read the file."""
    cache = os.environ.get('GG_BENCH_CACHE', '/tmp/gg_bench_cache')
    path = os.path.join(cache, f'generate_v2_lxd{lxd}_seed{seed}_G10_R2.pt')

    def build():
        from graingraphnn_b200 import generate as G
        hg = G.generate_graph(lxd=lxd, seed=seed, G=10.0, R=2.0, span=SPAN)
        x, ei, ea, geom = G.model_inputs(hg, lxd)
        d = {'x': x, 'ei': ei, 'ea': ea, 'glob': geom['global'], 'images': hg['tiling'].images, 'decimals': hg['tiling'].decimals,
             'offset': geom.get('domain_offset', 0), 'factor': geom['domain_factor']}
        try:
            os.makedirs(cache, exist_ok=True)
            torch.save(d, path + f'.tmp{os.getpid()}')
            os.replace(path + f'.tmp{os.getpid()}', path)
        except OSError:
            pass
        return d

    if world > 1:
        import torch.distributed as dist
        if rank == 0 and not os.path.exists(path):
            build()
        dist.barrier()
    d = torch.load(path) if os.path.exists(path) else build()
    ng, nj = d['x']['grain'].shape[0], d['x']['joint'].shape[0]
    ne = sum(int(v.shape[1]) for v in d['ei'].values())
    desc = (f'periodic hex-Voronoi grain domain of graph_trajectory.py --mode=generate (graingraphnn_b200/generate.py, validated '
            f'against the reference at lxd 40/120/240), lxd = {lxd} um = {lxd // 40} x {lxd // 40} patches, seed {seed}, G = 10, R = 2: '
            f'{ng} grains / {nj} joints / {ne} directed edges, grain in-degree 3..9')
    GEOM[lxd] = (d['offset'] if isinstance(d['offset'], torch.Tensor) else None, float(d['factor']))
    return d['x'], d['ei'], d['ea'], d['glob'], desc


def synth_weights():
    from graingraphnn_b200.weights import load_weights
    return load_weights(os.environ.get('GG_REGRESSOR_PT'), os.environ.get('GG_CLASSIFIER_PT'), head_gain=float(os.environ.get('GG_HEAD_GAIN', '0.02')))


def cpu_reference_run(steps, warmup, lxd=CPU_SAMPLE_LXD):
    """The reference-order CPU restatement (oracle/grain_oracle.py; PyG is not installable here) on a bounded sample of the
    workload: the same generator, seed and physical parameters on a smaller (or, time permitting, the same) domain, all host threads."""
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import grain_oracle as orc
    torch.set_num_threads(os.cpu_count() or 1)
    x, ei, ea, _, desc = make_domain(lxd)
    sd_r, sd_c, _ = synth_weights()
    edges = sum(int(v.shape[1]) for v in ei.values())
    x = {k: v.clone() for k, v in x.items()}

    def step(ea):
        """test.py:382-407, :562-575 in the reference's op order.  The grain-centre update of the GPU arm's step (test.py:471-476,
        :556-559 — a per-grain Python loop in the reference, 26 ms at 118 grains) is left out of the CPU arm: its
        cost there is the interpreter's, not the algorithm's, and leaving it out can only understate the ratio."""
        return orc.nn_step(sd_r, sd_c, x, ei, ea, SPAN)[1]

    with torch.no_grad():
        for _ in range(warmup):
            ea = step(ea)
        t0 = time.perf_counter()
        for _ in range(steps):
            ea = step(ea)
        dt = time.perf_counter() - t0
    return {'value': edges * steps / dt, 'unit': 'edges/s', 'cores': torch.get_num_threads(), 'kind': 'port',
            'sample': f'{steps} steps of the same generate-mode workload at lxd = {lxd} um ({x["grain"].shape[0]} grains, {edges} edges), '
                      f'oracle/grain_oracle.py in reference op order (without the per-grain Python loop of the grain-centre update)', 'ms_per_step': dt / steps * 1e3,
            'steps_per_s': steps / dt}


def reference_sample_lxd(steps, warmup, target_lxd, budget_s=180.0):
    """The largest domain of the ladder (up to the GPU arm's own) on which `warmup + steps` CPU steps are expected to end within
    `budget_s`: a 2-step probe at lxd = 240 gives edges/s, halved as a margin for the larger working set."""
    probe = cpu_reference_run(2, 1, lxd=CPU_SAMPLE_LXD)
    eps = 0.5 * probe['value']
    edges_240 = 75168.0
    for lxd in (1320, 920, 640, 480):
        if lxd <= target_lxd and (steps + warmup) * edges_240 * (lxd / 240.0) ** 2 / eps <= budget_s:
            return lxd
    return CPU_SAMPLE_LXD


def ncu_traffic(n_grains, launches):
    """dram__bytes_read.sum + dram__bytes_write.sum per gather launch (mean over the launches of one step) from the committed
    `ncu --set full` capture of this very workload (profiles/gather_traffic.json, written by scripts/ncu_traffic.py); None when
    the capture is of another size."""
    path = os.path.join(ROOT, 'profiles', 'gather_traffic.json')
    try:
        with open(path) as f:
            d = json.load(f)
        if d.get('n_grains') == n_grains and d.get('launches') == launches:
            return d['dram_bytes_per_launch']
    except (OSError, ValueError, KeyError):
        pass
    return None


def kernel_breakdown(eng, halo_times=None, reps=5):
    """Per-family device time of an eager step, CUDA events on the launching stream (torch's current stream) around every launch:
    `reps` steps enqueued back to back, per launch the MEDIAN over the steps (one step alone scatters by +-5 % from run to run)."""
    from graingraphnn_b200 import _lib
    times = {}

    class Timed:
        def __init__(self, inner):
            self._inner = inner

        def __getattr__(self, name):
            fn = getattr(self._inner, name)
            if name not in _lib.KERNELS_PER_CALL:          # queries (gg_version, gg_gather_tile_ecap, ...) launch nothing
                return fn

            def wrapped(*a):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                rc = fn(*a)
                e1.record()
                times.setdefault(name, []).append((e0, e1))
                return rc
            return wrapped

    from graingraphnn_b200 import engine as _engine
    real = _lib._LIB
    _lib._LIB = Timed(real)
    two = _engine._TWO_STREAMS
    _engine._TWO_STREAMS = False            # one stream: the events must bracket one kernel at a time
    real_exchange = None
    if halo_times is not None and getattr(eng, 'halo', None) is not None:
        real_exchange = eng.halo.exchange

        def timed_exchange(items):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            real_exchange(items)
            e1.record()
            halo_times.append((e0, e1))
        eng.halo.exchange = timed_exchange
    try:
        saved = eng._graph
        eng._graph = None
        torch.cuda._sleep(40_000_000)       # ~20 ms of device idle: the steps are enqueued behind it, so the events bracket
        for _ in range(reps):               # back-to-back kernels and not the host's launch cadence
            eng.step(SPAN)
        torch.cuda.synchronize()
        eng._graph = saved
    finally:
        _lib._LIB = real
        _engine._TWO_STREAMS = two
        if real_exchange is not None:
            eng.halo.exchange = real_exchange
    import numpy as np
    out = {}
    for k, v in times.items():
        per = len(v) // reps                                  # launches of this entry point per step (the same every step)
        ms = np.array([a.elapsed_time(b) for a, b in v[:per * reps]], dtype=np.float64).reshape(reps, per)
        each = np.median(ms, axis=0)
        out[k] = {'calls': per, 'ms_total': float(each.sum()), 'ms_each': [round(float(x), 4) for x in each]}
    if halo_times is not None and halo_times:
        per = len(halo_times) // reps
        del halo_times[:-per]                                 # the exchanges of the last of the steps
    return out


def build_engine(x, ei, ea, glob, dev, rank, world, lxd=None):
    """The engine of this rank with the graph resident and the geometry feedback of the reference's loop switched on (the grain
    centres follow the moved joints before the edge lengths are rebuilt, test.py:471-476, :556-575)."""
    sd_r, sd_c, _ = synth_weights()
    if world > 1:
        from graingraphnn_b200.partition import PartitionedEngine
        eng = PartitionedEngine.from_state_dicts(sd_r, sd_c, device=dev)
        eng.set_global_graph(x, ei, glob, rank, world, transport=os.environ.get('GG_HALO', 'auto'))
    else:
        from graingraphnn_b200.engine import RolloutEngine
        eng = RolloutEngine.from_state_dicts(sd_r, sd_c, device=dev)
        eng.set_graph({k: v.to(dev) for k, v in x.items()}, {k: v.to(dev) for k, v in ei.items()}, {k: v.to(dev) for k, v in ea.items()},
                      global_pos=None if os.environ.get('GG_BENCH_ORDER', 'morton') != 'morton' else glob)
    if lxd is not None and os.environ.get('GG_BENCH_FEEDBACK', '1') == '1':
        off, factor = GEOM[lxd]
        eng.enable_geometry_feedback(off, factor)
    return eng


def timed_steps(eng, steps, world, dev, barrier):
    """K steps between barriers, CUDA events on the launching stream, max over ranks -> ms for the K steps."""
    import torch.distributed as dist
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        eng.step(SPAN)
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def partition_parity(dev, rank, world, steps=3):
    """Before anything is timed: the slab-partitioned engine on `world` real ranks (this run's transport) against rank 0's
    undivided engine, on the reference-sized generate-mode domain (lxd = 240, 4,176 grains), `steps` rollout steps."""
    import numpy as np
    import torch.distributed as dist
    from graingraphnn_b200.engine import RolloutEngine
    x, ei, ea, glob, _ = make_domain(240, rank=rank, world=world)
    eng = build_engine(x, ei, ea, glob, dev, rank, world)      # (parity check without the feedback: three exchanges per step)
    outs = []
    for _ in range(steps):
        eng.step(SPAN)
        o = {k: (gid, v.cpu()) for k, (gid, v) in eng.owned_predictions().items()}
        o['x_joint'] = (eng.plan.own['joint'], eng.x['joint'][:eng.plan.n_own['joint']].cpu())
        outs.append(o)
    torch.cuda.synchronize()
    gathered = [None] * world
    dist.all_gather_object(gathered, outs)
    res = None
    if rank == 0:
        sd_r, sd_c, _ = synth_weights()
        single = RolloutEngine.from_state_dicts(sd_r, sd_c, dev)
        single.set_graph({k: v.to(dev) for k, v in x.items()}, {k: v.to(dev) for k, v in ei.items()})
        bit, worst = True, 0.0
        for s in range(steps):
            ref = {k: v.cpu() for k, v in single.step(SPAN).items()}
            ref['x_joint'] = single.x['joint'].cpu()
            for k in ('joint', 'grain', 'grain_area', 'edge_event', 'x_joint'):
                got = torch.full_like(ref[k], float('nan'))
                for r in range(world):
                    gid, val = gathered[r][s][k]
                    got[torch.from_numpy(np.asarray(gid))] = val
                if not torch.equal(got, ref[k]):
                    bit = False
                    worst = max(worst, float((got - ref[k]).abs().max() / ref[k].abs().max().clamp_min(1e-30)))
        res = {'graph': 'generate-mode lxd 240 (4176 grains)', 'steps': steps, 'ranks': world, 'bit_identical': bit, 'max_rel': worst,
               'ok': bool(bit or worst < 1e-5), 'transport': eng.halo.transport,
               'note': 'a row whose in-edges straddle a gather tile is summed in other softmax chunks on a slab than on the undivided graph'}
    del eng
    torch.cuda.empty_cache()
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--lxd', type=int, default=None, help='domain edge in um, a multiple of 40 (default: ~1.25e5 grains per GPU)')
    ap.add_argument('--no-graph', action='store_true', help='do not replay the step from a CUDA graph')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-strong', action='store_true', help='skip the strong-scaling measurement on the 10^6-grain domain')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'ours' else args.warmup
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    lxd = args.lxd or weak_lxd(world if args.impl == 'ours' else max(args.gpus, 1))
    step_desc = ('nn-step: regressor+classifier fwd (enc+dec HeteroPGCLSTM), heads, feature update, grain centres from the moved joints, '
                 'edge-length rebuild (test.py:382-407, :471-476, :556-575); fixed topology')

    if args.impl == 'reference':
        if rank != 0:
            return 0
        steps, warmup = args.steps, args.warmup
        r = cpu_reference_run(steps, warmup, lxd=reference_sample_lxd(steps, warmup, lxd))
        print(json.dumps({'metric': 'rollout_edges_per_sec', 'value': r['value'], 'unit': 'edges/s', 'impl': 'reference',
                          'n_gpus': args.gpus, 'steps': steps, 'warmup': warmup, 'ms_per_step': r['ms_per_step'],
                          'steps_per_sec': r['steps_per_s'], 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
                          'dtype': 'f32', 'data': 'synthetic',
                          'config': {'workload': f'generate-mode periodic grain domain (lxd = {lxd} um in the GPU arm); reference-order CPU '
                                                 f'restatement (PyG is not installable here) timed on a bounded sample of it',
                                     'step': step_desc, 'sample': r['sample']},
                          'cpu_baseline': {k: r[k] for k in ('value', 'unit', 'cores', 'kind', 'sample')},
                          'e2e': {'value': r['value'], 'unit': 'edges/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}))
        return 0

    assert torch.cuda.is_available(), 'bench.py needs a CUDA device (there is no CPU fallback)'
    import torch.distributed as dist
    from graingraphnn_b200 import _lib
    dev = torch.device('cuda', local_rank)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    n_gpus = world
    pk = peaks()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    parity = partition_parity(dev, rank, world) if world > 1 else None

    x, ei, ea, glob, desc = make_domain(lxd, rank=rank, world=world)
    eng = build_engine(x, ei, ea, glob, dev, rank, world, lxd)
    ng_total, nj_total = x['grain'].shape[0], x['joint'].shape[0]
    edges_total = sum(int(v.shape[1]) for v in ei.values())
    weights_desc = synth_weights()[2]

    use_graph = not args.no_graph and (world == 1 or eng.halo.transport == 'p2p')
    for _ in range(args.warmup):
        eng.step(SPAN)
    # per-family times of one eager step BEFORE the sustained load of the timed regions: SM clock at its maximum, no power cap yet
    bd_cool = kernel_breakdown(eng) if world == 1 else None
    if use_graph:
        eng.capture(SPAN, warmup=1)
    barrier()

    # ---- timed region: device-resident inputs ---------------------------------------------------------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = _lib.LAUNCHES[0]
    ms = timed_steps(eng, args.steps, world, dev, barrier)
    launches = _lib.LAUNCHES[0] - l0
    if use_graph:
        launches = eng.launches_per_step * args.steps
    clocks = sampler.stop() if rank == 0 else None

    # ---- end-to-end: host buffers in, host predictions out, through the public engine API -------------------
    hx = {k: v.clone().pin_memory() for k, v in eng.host_features().items()}
    hout = None
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    for _ in range(args.steps):
        eng.load_features(hx)                    # H2D of this step's node features (pinned)
        pred = eng.step(SPAN)
        hout = eng.fetch_predictions(pred, hout)  # D2H of joint / grain / grain_area / edge_event (pinned)
    e3.record()
    barrier()
    t = torch.tensor([e2.elapsed_time(e3)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_e2e = float(t.item())
    h2d = sum(v.numel() * 4 for v in hx.values())
    d2h = sum(v.numel() * 4 for v in hout.values())
    if world > 1:
        tt = torch.tensor([h2d, d2h], dtype=torch.float64, device=dev)
        dist.all_reduce(tt)
        h2d, d2h = int(tt[0].item()), int(tt[1].item())

    # ---- per-family device time of one eager step (all ranks take part: the step contains the halo exchanges) ----
    halo_ev = []
    bd = kernel_breakdown(eng, halo_ev)
    halo = None
    if world > 1:
        c = eng.counts()
        hms = torch.tensor([sum(a.elapsed_time(b) for a, b in halo_ev), float(sum(eng.halo.bytes_sent_per_exchange[-len(halo_ev):])) if halo_ev else 0.0,
                            float(c['halo_grain'] + c['halo_joint'])], dtype=torch.float64, device=dev)
        mx = hms.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        dist.all_reduce(hms)
        halo = {'exchanges_per_step': len(halo_ev), 'bytes_per_step': int(hms[1].item()), 'ms_per_step': float(mx[0].item()),
                'halo_rows_total': int(hms[2].item()), 'transport': eng.halo.transport,
                'note': 'ms = pack kernels + transfers + waits of the slowest rank in one eager step (not overlapped with compute); bytes summed over ranks'}
    barrier()

    # ---- strong scaling: the fixed 10^6-grain domain of BASELINE config 4 on this run's N GPUs ----------------
    strong = None
    alg = eng.algorithmic_work()
    n_grain_local = eng.counts()['n_grain']
    if not args.no_strong and os.environ.get('GG_BENCH_STRONG', '1') != '0':
        if lxd == STRONG_LXD:
            strong = {'lxd': lxd, 'grains': ng_total, 'n_gpus': n_gpus, 'ms_per_step': ms / args.steps, 'steps_per_sec': args.steps / (ms / 1e3),
                      'steps': args.steps, 'same_run_as': 'value'}
        else:
            try:
                del hx, hout
                engs = None
                sx, sei, sea, sglob, _ = make_domain(STRONG_LXD, rank=rank, world=world)
                engs = build_engine(sx, sei, sea, sglob, dev, rank, world, STRONG_LXD)
                ksteps = max(3, min(args.steps, 8))
                for _ in range(3):
                    engs.step(SPAN)
                if use_graph:
                    engs.capture(SPAN, warmup=1)
                sms = timed_steps(engs, ksteps, world, dev, barrier)
                strong = {'lxd': STRONG_LXD, 'grains': int(sx['grain'].shape[0]), 'n_gpus': n_gpus, 'ms_per_step': sms / ksteps,
                          'steps_per_sec': ksteps / (sms / 1e3), 'steps': ksteps, 'warmup': 3}
                del engs
                torch.cuda.empty_cache()
            except Exception as exc:                                    # never lose the headline line to the extra measurement
                strong = {'error': repr(exc)[:300]}
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel (rank 0, live CUDA events) -----------------------------------------
    for name in ('gg_pgat_gather_tiled', 'gg_pgat_gather_tiled_multi'):      # all entry points of kernel family (b)
        if name in bd:
            t = bd.pop(name)
            g = bd.setdefault('gg_pgat_gather', {'calls': 0, 'ms_total': 0.0, 'ms_each': []})
            g['calls'] += t['calls']; g['ms_total'] += t['ms_total']; g['ms_each'] += t['ms_each']
    fam = {k: v['ms_total'] for k, v in bd.items()}
    total_fam = sum(fam.values())
    top = max(fam, key=fam.get)
    work = {'gg_pgat_gather': ('hbm', alg['gg_pgat_gather'], 'GB/s', pk['hbm_gbs']),
            'gg_node_proj': ('tensor', alg['gg_node_proj'], 'TFLOP/s', pk['bf16_sustained']),
            'gg_node_proj_tc': ('tensor', alg['gg_node_proj'], 'TFLOP/s', pk['bf16_sustained']),
            'gg_node_proj_fused': ('tensor', alg['gg_node_proj'], 'TFLOP/s', pk['bf16_sustained']),
            'gg_gate_update': ('tensor', alg['gg_gate_update'], 'TFLOP/s', pk['bf16_sustained']),
            'gg_gate_update_tc': ('tensor', alg['gg_gate_update'], 'TFLOP/s', pk['bf16_sustained'])}
    if top in work:
        bound, amount, unit, peak = work[top]
        dur_s = fam[top] / 1e3
        achieved = amount / dur_s / (1e9 if unit == 'GB/s' else 1e12)
        roof = {'kernel': top, 'bound': bound, 'achieved': achieved, 'peak': peak, 'unit': unit, 'frac': achieved / peak,
                'traffic': None, 'peak_source': pk['source'] + (' (sustained bf16 cuBLAS)' if bound == 'tensor' else ' (copy)'),
                'launches_per_step': bd[top]['calls'], 'ms_per_step': fam[top], 'share_of_step': fam[top] / total_fam}
        if top == 'gg_pgat_gather':
            # per-launch figures like `traffic`: algorithmic bytes and time of the average launch of the step
            roof['achieved_bytes_per_launch'] = amount / bd[top]['calls']
            roof['traffic'] = ncu_traffic(n_grain_local, bd[top]['calls'])
            roof['launch'] = 'one launch per cell: its three edge types are segments of the same persistent kernel'
            if bd_cool is not None:
                cool = sum(v['ms_total'] for k, v in bd_cool.items() if k.startswith('gg_pgat_gather'))
                roof['before_sustained_load'] = {
                    'ms_per_step': cool, 'achieved': amount / (cool / 1e3) / 1e9, 'frac': amount / (cool / 1e3) / 1e9 / peak,
                    'note': 'the same measurement on the same step right after the warm-up steps; `frac` above is taken after the timed '
                            'regions, when the board sits at its 1 kW cap (sw_power_cap) and the SM clock has dropped: the gather is bound by '
                            'the consumer warps\' instruction latency, so it follows the SM clock, a plain device copy does not '
                            '(profiles/r2_hot_cold.txt)'}
    else:
        roof = {'kernel': top, 'bound': 'hbm', 'achieved': None, 'peak': pk['hbm_gbs'], 'unit': 'GB/s', 'frac': None, 'traffic': None}
    roof['breakdown_ms'] = {k: round(v, 4) for k, v in sorted(fam.items(), key=lambda kv: -kv[1])}
    if os.environ.get('GG_BENCH_VERBOSE'):
        roof['per_call_ms'] = {k: v['ms_each'] for k, v in bd.items()}

    # ---- the same step with the rows SURVEY §8f adds switched on (not part of `value`): grain centres from the moved
    #      joints (gg_region_center, row f2) and the event candidates of the host topology update (gg_select_events, row f1)
    widened = None
    if world == 1:
        try:
            eng.enable_event_selection()
            if use_graph:
                eng.capture(SPAN, warmup=2)
            else:
                eng.step(SPAN)
            torch.cuda.synchronize()
            e4, e5 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e4.record()
            for _ in range(args.steps):
                eng.step(SPAN)
            e5.record()
            torch.cuda.synchronize()
            ev = eng.fetch_events()
            widened = {'ms_per_step': e4.elapsed_time(e5) / args.steps, 'launches_per_step': eng.launches_per_step,
                       'adds': 'gg_select_events (models.py:627-629, test.py:414): the event candidates of the host topology update stay on the device',
                       'event_candidates_last_step': [int(ev['L1'].numel()), int(ev['grain_event'].numel())],
                       'event_d2h_bytes': 8 + 8 * int(ev['L1'].numel() + ev['grain_event'].numel())}
        except Exception as exc:                                   # never lose the headline line to the extra measurement
            widened = {'error': repr(exc)[:200]}

    cpu = None if args.no_cpu_baseline else cpu_reference_run(5, 1, lxd=CPU_BASELINE_LXD)
    steps_per_s = args.steps / (ms / 1e3)
    part = (f'; x-slab partition over {n_gpus} GPUs, {halo["transport"]} halo exchange per message-passing hop' if n_gpus > 1 else '; single B200 rollout')
    line = {
        'metric': 'rollout_edges_per_sec', 'value': edges_total * steps_per_s, 'unit': 'edges/s', 'n_gpus': n_gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms / args.steps, 'steps_per_sec': steps_per_s,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': desc + part, 'step': step_desc, 'weights': weights_desc,
                   'l2': 'per-step working set (projections, GBs) exceeds the 126 MB L2; no explicit flush',
                   'rows': 'the engine keeps its node rows along a Morton curve of the global positions (RolloutEngine.set_graph(global_pos=...)); inputs, '
                           'predictions and events keep the generator\'s numbering',
                   'cuda_graph': bool(use_graph), 'gemm': os.environ.get('GG_GEMM', 'auto'),
                   'streams': 'regressor and classifier cells on two streams (two graph branches); per-kernel times from a one-stream step'},
        'clocks': clocks,
        'e2e': {'value': edges_total * args.steps / (ms_e2e / 1e3), 'unit': 'edges/s', 'h2d_bytes_per_step': h2d,
                'd2h_bytes_per_step': d2h, 'ms_per_step': ms_e2e / args.steps},
        'gpu_launches': int(launches),
        'roofline': roof,
        'step_with_event_selection': widened,
        'cpu_baseline': None if cpu is None else {k: cpu[k] for k in ('value', 'unit', 'cores', 'kind', 'sample')},
    }
    if n_gpus > 1:
        line['partition_parity'] = parity
        line['halo'] = halo
    if strong is not None:
        line['strong_scaling'] = strong
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == '__main__':
    sys.exit(main())
