#!/bin/bash
# One gpurun call: GPU tests, bench (N=1), ncu launch list and one full capture of the gather kernel.
set -x
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
tail -c 3000 gpurun_out/bench_n1.json
GG_BENCH_VERBOSE=1 python bench.py --steps 5 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/bench_n1_eager.json 2>> gpurun_out/bench_n1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pgat_gather -s 24 -c 3 -o gpurun_out/prof_gather \
    python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu_gather.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'node_proj_tc|gate_update_tc' -s 16 -c 4 -o gpurun_out/prof_gemm \
    python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu_gemm.log 2>&1
ls -la gpurun_out
