#!/usr/bin/env python
"""BASELINE config 5: an ensemble of 64 independent 40x40 rollouts over a (G, R) sweep, one block-diagonal batch per GPU, no
communication (replicas only — DESIGN.md §6).

    python scripts/ensemble_bench.py                       # 1 GPU: the 64 rollouts as 8 batches of 8, one after the other
    GG_ENSEMBLE_BATCH=64 python scripts/ensemble_bench.py  # 1 GPU: all 64 in one block-diagonal batch
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 scripts/ensemble_bench.py    # 8 rollouts per GPU

Every member is what `graph_trajectory.py --mode=generate --lxd 40 --seed s --G g --R r` builds (graingraphnn_b200/generate.py,
seeds 1..64, (G, R) on the 8 x 8 grid over [0.5, 10] x [0.2, 2]) with the span the reference's nearest-neighbour lookup in
GR_train_grid.pkl gives for its (G, R) (graph_trajectory.py:1308-1316; the table below was produced with
generate.span_from_grid on the reference's file, which is not on the GPU box).  A member with span s takes (120 // s) steps
(test.py:353 `range(span, frames, span)`, frames = 121); a batch steps until its slowest member (smallest span) is done, every
member advancing by its own span (EnsembleEngine.step(spans)); finished members idle at the z clamp (test.py:405-407).
Prints one JSON line: rollouts/s and member-steps/s over the whole ensemble, max over ranks of the device time."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from graingraphnn_b200 import generate as G  # noqa: E402
from graingraphnn_b200.ensemble import EnsembleEngine  # noqa: E402
from graingraphnn_b200.weights import load_weights  # noqa: E402

SPANS = [[6, 6, 6, 6, 6, 6, 6, 6], [12, 6, 6, 6, 6, 6, 6, 6], [24, 8, 6, 6, 6, 6, 6, 6], [60, 10, 8, 6, 6, 6, 6, 6],
         [60, 12, 8, 8, 6, 6, 6, 6], [120, 20, 8, 8, 6, 6, 6, 6], [60, 30, 10, 8, 8, 6, 6, 6], [60, 40, 12, 6, 6, 6, 6, 6]]
GS, RS = np.linspace(0.5, 10.0, 8), np.linspace(0.2, 2.0, 8)


def member(i):
    gi, ri = divmod(i, 8)
    hg = G.generate_graph(lxd=40, seed=i + 1, G=float(GS[gi]), R=float(RS[ri]), span=SPANS[gi][ri])
    x, ei, ea, _ = G.model_inputs(hg, 40)
    return (x, ei, ea), SPANS[gi][ri]


def main():
    rank, world, local = int(os.environ.get('RANK', '0')), int(os.environ.get('WORLD_SIZE', '1')), int(os.environ.get('LOCAL_RANK', '0'))
    dev = torch.device('cuda', local)
    torch.cuda.set_device(dev)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=dev)
    per_batch = int(os.environ.get('GG_ENSEMBLE_BATCH', '8'))                     # 8 = BASELINE config 5's "8 per GPU"; 64 = the whole ensemble as one block-diagonal batch
    assert 64 % per_batch == 0
    mine = [b for b in range(64 // per_batch) if b % world == rank]               # batches of this rank
    sd_r, sd_c, wdesc = load_weights(os.environ.get('GG_REGRESSOR_PT'), os.environ.get('GG_CLASSIFIER_PT'))
    eng = EnsembleEngine.from_state_dicts(sd_r, sd_c, device=dev)
    batches = []
    for b in mine:
        ms = [member(b * per_batch + k) for k in range(per_batch)]
        batches.append(([m[0] for m in ms], tuple(m[1] for m in ms)))
    total_member_steps = sum(120 // s for _, spans in batches for s in spans)
    ms_total, launches = 0.0, 0
    for rep in range(2):                                                           # rep 0 warms up (lazy tile indices, allocator)
        ms_total = 0.0
        for graphs, spans in batches:
            eng.set_graphs([({k: v.to(dev) for k, v in g[0].items()}, {k: v.to(dev) for k, v in g[1].items()},
                             {k: v.to(dev) for k, v in g[2].items()}) for g in graphs])
            n_steps = 120 // min(spans)
            eng.capture(spans, warmup=1)                                           # fixed topology per batch: replay the step graph
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(n_steps):
                eng.step(spans)
            e1.record()
            torch.cuda.synchronize()
            ms_total += e0.elapsed_time(e1)
            launches += eng.launches_per_step * n_steps
    t = torch.tensor([ms_total, float(total_member_steps)], dtype=torch.float64, device=dev)
    if world > 1:
        import torch.distributed as dist
        mx = t.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        dist.all_reduce(t)
        ms_total, total_member_steps = float(mx[0].item()), int(t[1].item())
    if rank == 0:
        c = eng.counts()
        print(json.dumps({'metric': 'ensemble_rollouts_per_sec', 'value': 64 / (ms_total / 1e3), 'unit': 'rollouts/s', 'n_gpus': world,
                          'member_steps_per_sec': total_member_steps / (ms_total / 1e3), 'ms_total': ms_total,
                          'config': {'workload': '64 generate-mode 40x40 rollouts (seeds 1..64, (G, R) on an 8 x 8 grid, spans 6..120 by the '
                                                 'reference\'s lookup), block-diagonal batches of ' + str(per_batch) + ', one batch at a time per GPU, no communication',
                                     'batch_nodes': c, 'weights': wdesc, 'cuda_graph': True,
                                     'step': 'nn-step per member with its own span; fixed topology'}}))
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
