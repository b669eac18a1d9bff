#!/bin/bash
# One gpurun call: GPU tests, bench (N=1), optionally ncu launch list and full captures of the hot kernels.
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x --timeout 120 2>&1 | tail -15
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
tail -c 2500 gpurun_out/bench_n1.json; tail -5 gpurun_out/bench_n1.err
if [ -n "$GG_AB" ]; then
GG_GATHER=items timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1_items.json 2>> gpurun_out/bench_n1.err
tail -c 900 gpurun_out/bench_n1_items.json
fi
if [ -n "$GG_PROFILE" ]; then
KREG='regex:pgat_gather|node_proj|gate_update|split_tf32|edge_length|edge_head|node_head|feature_update|z_probe|z_clamp|permute_kernel'
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREG" -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
# the 12 gather launches of ONE step (3 warm-up steps x 12 launches are skipped), full metric set + source
timeout 800 ncu --set full --clock-control none --import-source on -k regex:pgat_gather -s 36 -c 12 -o gpurun_out/prof_gather_step \
    python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu_gather.log 2>&1
fi
ls -la gpurun_out
