"""Multi-GPU check, launched by tests/test_partition.py::test_multi_gpu_partition (or by hand):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tests/mgpu_check.py
Every rank steps its slab with real halo exchanges (GG_HALO = nccl | p2p); rank 0 also steps the undivided graph on its own
GPU and checks that the union of the owned outputs is identical."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tests')):
    if p not in sys.path:
        sys.path.insert(0, p)

import grain_oracle as orc  # noqa: E402
from graingraphnn_b200.engine import RolloutEngine  # noqa: E402
from graingraphnn_b200.partition import PartitionedEngine  # noqa: E402
from graingraphnn_b200.synth import honeycomb_graph, lattice_dims  # noqa: E402


def main():
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    dev = torch.device('cuda', local)
    torch.cuda.set_device(dev)
    dist.init_process_group('nccl', device_id=dev)
    px, py = 4 * world, 3
    nx, ny = lattice_dims(px, py)
    x, ei, glob = honeycomb_graph(nx, ny, seed=11, patches=(px, py), return_global=True)
    sd_r, sd_c = orc.synth_state_dict('regressor', 1), orc.synth_state_dict('classifier', 2)
    eng = PartitionedEngine.from_state_dicts(sd_r, sd_c, device=dev)
    eng.set_global_graph(x, ei, glob, rank, world)
    feedback = os.environ.get('GG_FEEDBACK', '0') == '1'      # row f2: grain centres follow the joints (fourth exchange)
    if feedback:
        eng.enable_geometry_feedback()
    steps = 3
    outs = []
    for _ in range(steps):
        eng.step(6)
        o = {k: (gid, v.cpu()) for k, (gid, v) in eng.owned_predictions().items()}
        o['x_grain'] = (eng.plan.own['grain'], eng.x['grain'][:eng.plan.n_own['grain']].cpu())
        outs.append(o)
    torch.cuda.synchronize()
    gathered = [None] * world
    dist.all_gather_object(gathered, outs)
    ok = True
    if rank == 0:
        single = RolloutEngine.from_state_dicts(sd_r, sd_c, dev)
        single.set_graph({k: v.to(dev) for k, v in x.items()}, {k: v.to(dev) for k, v in ei.items()})
        if feedback:
            single.enable_geometry_feedback()
        for s in range(steps):
            ref = {k: v.cpu() for k, v in single.step(6).items()}
            ref['x_grain'] = single.x['grain'].cpu()
            for k in ('joint', 'grain', 'grain_area', 'edge_event', 'x_grain'):
                got = torch.full_like(ref[k], float('nan'))
                for r in range(world):
                    gid, val = gathered[r][s][k]
                    got[torch.from_numpy(np.asarray(gid))] = val
                same = bool(torch.equal(got, ref[k]))
                if not same:
                    # Rows whose in-edges straddle a tile boundary of the gather are summed in different chunks on the slab and
                    # on the undivided graph (the tiling follows the CSR position): equal to rounding, not bit for bit.
                    err = float((got - ref[k]).abs().max() / ref[k].abs().max().clamp_min(1e-30))
                    same = err < 1e-5
                    print(f'step {s} {k}: not bit-identical, max rel err {err:.3e} ({"within" if same else "OUTSIDE"} 1e-5)', flush=True)
                ok &= same
        print(f'MGPU {"OK" if ok else "FAIL"} world={world} transport={eng.halo.transport} feedback={int(feedback)} '
              f'halo_bytes_per_exchange={eng.halo.bytes_sent_per_exchange[:4]} counts={eng.counts()}', flush=True)
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    return 0 if int(flag.item()) == 1 else 1


if __name__ == '__main__':
    sys.exit(main())
