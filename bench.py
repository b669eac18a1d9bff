#!/usr/bin/env python
"""bench.py — GrainGNN rollout throughput (edges/s, steps/s) on B200, with roofline and CPU baseline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--patches PXxPY]

Workload (weak scaling): a synthetic periodic hexagonal-lattice grain domain of 36 x 30 patches (40 um each) PER GPU,
slabs side by side along x: N=1 -> 124,560 grains / 2.24 M directed edges (the ~10^5-grain single-B200 config),
N=8 -> 288 x 30 patches = 996,480 grains / 17.9 M edges (the ~10^6-grain slab-partitioned config).
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
for p in (ROOT, os.path.join(ROOT, 'oracle')):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

PATCHES_PER_GPU = (36, 30)
SPAN = 6


def peaks():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            d = json.load(f)
        return {'hbm_gbs': d['hbm_gbs'], 'bf16_tflops': d['bf16_tflops'], 'bf16_sustained': d.get('bf16_tflops_sustained', d['bf16_tflops']),
                'source': 'measured'}
    except Exception:
        return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'bf16_sustained': 1400.0, 'source': 'fallback'}


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled every 200 ms while the timed region runs."""
    Q = 'index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,' \
        'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits', '-lms', '200',
                                          '-i', str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
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


def make_domain(n_gpus, patches=None, seed=1):
    from graingraphnn_b200.synth import honeycomb_graph, lattice_dims
    px, py = patches if patches else (PATCHES_PER_GPU[0] * n_gpus, PATCHES_PER_GPU[1])
    nx, ny = lattice_dims(px, py)
    x, ei, glob = honeycomb_graph(nx, ny, seed=seed, patches=(px, py), return_global=True)
    return x, ei, glob, (px, py)


def synth_weights():
    import grain_oracle as orc   # weights only (seeded stand-ins: the shipped .pt files are absent); not on the timed path
    return orc.synth_state_dict('regressor', 1), orc.synth_state_dict('classifier', 2)


def cpu_reference_run(steps, warmup, sample_patches=(3, 3)):
    """The reference-order CPU restatement (oracle) on a bounded sample: a 3 x 3 patch domain (~1,040 grains, the size of
    the reference's 120x120 case), all host threads."""
    import grain_oracle as orc
    torch.set_num_threads(os.cpu_count() or 1)
    x, ei, _, pp = make_domain(1, sample_patches, seed=1)
    ea = orc.edge_attr_rebuild(x, ei)
    sd_r, sd_c = synth_weights()
    edges = sum(int(v.shape[1]) for v in ei.values())
    with torch.no_grad():
        for _ in range(warmup):
            _, ea = orc.nn_step(sd_r, sd_c, x, ei, ea, SPAN)
        t0 = time.perf_counter()
        for _ in range(steps):
            _, ea = orc.nn_step(sd_r, sd_c, x, ei, ea, SPAN)
        dt = time.perf_counter() - t0
    return {'value': edges * steps / dt, 'unit': 'edges/s', 'cores': torch.get_num_threads(), 'kind': 'port',
            'sample': f'{steps} steps of a {pp[0]}x{pp[1]}-patch synthetic domain ({x["grain"].shape[0]} grains, {edges} edges), '
                      f'oracle/grain_oracle.py in reference op order', 'ms_per_step': dt / steps * 1e3,
            'steps_per_s': steps / dt}


def ncu_traffic(n_grains, launches):
    """dram__bytes_read.sum + dram__bytes_write.sum per gather launch (mean over the launches of one step) from the committed
    `ncu --set full` capture of this very workload (profiles/gather_traffic.json, written by scripts/ncu_traffic.py); None when
    the capture is of another size."""
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'profiles', 'gather_traffic.json')
    try:
        with open(path) as f:
            d = json.load(f)
        if d.get('n_grains') == n_grains and d.get('launches') == launches:
            return d['dram_bytes_per_launch']
    except (OSError, ValueError, KeyError):
        pass
    return None


def kernel_breakdown(eng, peak):
    """Per-family device time of ONE eager step, CUDA events on the launching stream (torch's current stream)."""
    from graingraphnn_b200 import _lib, cell, heads, graph
    times = {}
    L = _lib.lib()

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
    try:
        saved = eng._graph
        eng._graph = None
        torch.cuda._sleep(40_000_000)       # ~20 ms of device idle: the whole step is enqueued behind it, so the events bracket
        eng.step(SPAN)                      # back-to-back kernels and not the host's launch cadence
        torch.cuda.synchronize()
        eng._graph = saved
    finally:
        _lib._LIB = real
        _engine._TWO_STREAMS = two
    return {k: {'calls': len(v), 'ms_total': sum(a.elapsed_time(b) for a, b in v), 'ms_each': [round(a.elapsed_time(b), 4) for a, b in v]}
            for k, v in times.items()}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--patches', default=None, help='PXxPY total domain in 40-um patches (default 36N x 30)')
    ap.add_argument('--no-graph', action='store_true', help='do not replay the step from a CUDA graph')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'ours' else args.warmup
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    patches = tuple(int(v) for v in args.patches.lower().split('x')) if args.patches else None

    if args.impl == 'reference':
        if rank != 0:
            return 0
        steps, warmup = min(args.steps, 10), min(args.warmup, 2)
        r = cpu_reference_run(steps, warmup)
        print(json.dumps({'metric': 'rollout_edges_per_sec', 'value': r['value'], 'unit': 'edges/s', 'impl': 'reference',
                          'n_gpus': args.gpus, 'steps': steps, 'warmup': warmup, 'ms_per_step': r['ms_per_step'],
                          'steps_per_sec': r['steps_per_s'], 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
                          'dtype': 'f32', 'data': 'synthetic',
                          'config': {'workload': 'GrainGNN nn-step (regressor+classifier fwd, heads, feature update, edge-length '
                                                 'rebuild), reference-order CPU restatement; PyG is not installable here',
                                     'sample': r['sample']},
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

    x, ei, glob, pp = make_domain(n_gpus, patches)
    sd_r, sd_c = synth_weights()
    if world > 1:
        from graingraphnn_b200.partition import PartitionedEngine
        eng = PartitionedEngine.from_state_dicts(sd_r, sd_c, device=dev)
        eng.set_global_graph(x, ei, glob, rank, world)
    else:
        from graingraphnn_b200.engine import RolloutEngine
        eng = RolloutEngine.from_state_dicts(sd_r, sd_c, device=dev)
        eng.set_graph({k: v.to(dev) for k, v in x.items()}, {k: v.to(dev) for k, v in ei.items()})
    ng_total, nj_total = x['grain'].shape[0], x['joint'].shape[0]
    edges_total = sum(int(v.shape[1]) for v in ei.values())
    cnt = eng.counts()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    use_graph = not args.no_graph and world == 1
    for _ in range(args.warmup):
        eng.step(SPAN)
    if use_graph:
        eng.capture(SPAN, warmup=1)
    barrier()

    # ---- timed region: device-resident inputs ---------------------------------------------------------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = _lib.LAUNCHES[0]
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        eng.step(SPAN)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = _lib.LAUNCHES[0] - l0
    if use_graph:
        launches = eng.launches_per_step * args.steps
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
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
    bd = kernel_breakdown(eng, pk)
    barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel (rank 0, live CUDA events) -----------------------------------------
    if 'gg_pgat_gather_tiled' in bd:                       # both entry points are kernel family (b)
        t = bd.pop('gg_pgat_gather_tiled')
        g = bd.setdefault('gg_pgat_gather', {'calls': 0, 'ms_total': 0.0, 'ms_each': []})
        g['calls'] += t['calls']; g['ms_total'] += t['ms_total']; g['ms_each'] += t['ms_each']
    fam = {k: v['ms_total'] for k, v in bd.items()}
    total_fam = sum(fam.values())
    top = max(fam, key=fam.get)
    alg = eng.algorithmic_work()
    work = {'gg_pgat_gather': ('hbm', alg['gg_pgat_gather'], 'GB/s', pk['hbm_gbs']),
            'gg_node_proj': ('tensor', alg['gg_node_proj'], 'TFLOP/s', pk['bf16_sustained']),
            'gg_node_proj_tc': ('tensor', alg['gg_node_proj'], 'TFLOP/s', pk['bf16_sustained']),
            'gg_node_proj_fused': ('tensor', alg['gg_node_proj'], 'TFLOP/s', pk['bf16_sustained']),
            'gg_gate_update': ('tensor', alg['gg_gate_update'], 'TFLOP/s', pk['bf16_sustained'])}
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
            roof['traffic'] = ncu_traffic(ng_total, bd[top]['calls'])
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
            eng.enable_geometry_feedback()
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
                       'adds': 'gg_region_center (grain centres, test.py:476 + :556-559) and gg_select_events (models.py:627-629, test.py:414)',
                       'event_candidates_last_step': [int(ev['L1'].numel()), int(ev['grain_event'].numel())],
                       'event_d2h_bytes': 8 + 8 * int(ev['L1'].numel() + ev['grain_event'].numel())}
        except Exception as exc:                                   # never lose the headline line to the extra measurement
            widened = {'error': repr(exc)[:200]}

    cpu = None if args.no_cpu_baseline else cpu_reference_run(5, 1)
    steps_per_s = args.steps / (ms / 1e3)
    line = {
        'metric': 'rollout_edges_per_sec', 'value': edges_total * steps_per_s, 'unit': 'edges/s', 'n_gpus': n_gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms / args.steps, 'steps_per_sec': steps_per_s,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': f'synthetic periodic hex-lattice grain domain {pp[0]}x{pp[1]} patches (40 um), {ng_total} grains / '
                               f'{nj_total} joints / {edges_total} directed edges, {PATCHES_PER_GPU[0]}x{PATCHES_PER_GPU[1]} patches per GPU, '
                               f'x-slab partition + NCCL halo exchange' if n_gpus > 1 else
                               f'synthetic periodic hex-lattice grain domain {pp[0]}x{pp[1]} patches (40 um), {ng_total} grains / '
                               f'{nj_total} joints / {edges_total} directed edges, single B200 rollout',
                   'step': 'nn-step: regressor+classifier fwd (enc+dec HeteroPGCLSTM), heads, feature update, edge-length rebuild; fixed topology',
                   'weights': 'seeded stand-ins with the reference state_dict layout (regressor0.pt/classifier1.pt absent)',
                   'l2': 'per-step working set (projections, GBs) exceeds the 126 MB L2; no explicit flush',
                   'cuda_graph': bool(use_graph), 'gemm': os.environ.get('GG_GEMM', 'auto'),
                   'streams': 'regressor and classifier cells on two streams (two graph branches); per-kernel times from a one-stream step'},
        'clocks': clocks,
        'e2e': {'value': edges_total * args.steps / (ms_e2e / 1e3), 'unit': 'edges/s', 'h2d_bytes_per_step': h2d,
                'd2h_bytes_per_step': d2h, 'ms_per_step': ms_e2e / args.steps},
        'gpu_launches': int(launches),
        'roofline': roof,
        'step_with_geometry_feedback_and_event_selection': widened,
        'cpu_baseline': None if cpu is None else {k: cpu[k] for k in ('value', 'unit', 'cores', 'kind', 'sample')},
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == '__main__':
    sys.exit(main())
