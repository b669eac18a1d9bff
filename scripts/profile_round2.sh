#!/bin/bash
# One gpurun call (1 GPU): the ncu evidence of round 2 on the bench workload (bench.py, N = 1).
#   gpurun_out/r2_launches.csv        every launch of two eager steps with its device time (shares, not absolutes)
#   gpurun_out/r2_prof_gather.ncu-rep `--set full` of the 4 gather launches of one step (one per cell)
#   gpurun_out/r2_prof_gather_split.ncu-rep  the same step with one launch per edge type (GG_GATHER_MERGE=0): 12 launches
#   gpurun_out/r2_prof_gemm.ncu-rep   `--set full` of the 16 GEMM launches of one step
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-strong"
KREG='regex:pgat_gather|node_proj|gate_update|edge_refresh|edge_head|node_head|feature_update|z_probe|z_clamp|region_center|index_select|indexSelect'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREG" -s 120 -c 80 --csv --log-file gpurun_out/r2_launches.csv $B > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:pgat_gather -s 12 -c 4 -f -o gpurun_out/r2_prof_gather $B > gpurun_out/ncu_gather.log 2>&1
GG_GATHER_MERGE=0 timeout 900 ncu --set full --clock-control none -k regex:pgat_gather -s 36 -c 12 -f -o gpurun_out/r2_prof_gather_split $B > gpurun_out/ncu_gather_split.log 2>&1
timeout 1200 ncu --set full --clock-control none -k 'regex:node_proj|gate_update' -s 48 -c 16 -f -o gpurun_out/r2_prof_gemm $B > gpurun_out/ncu_gemm.log 2>&1
# the reports are large (gpurun brings back <= 64 MiB): keep the gather report, turn the others into raw CSV pages on the box
for r in r2_prof_gather r2_prof_gather_split r2_prof_gemm; do ncu -i gpurun_out/$r.ncu-rep --page raw --csv > gpurun_out/$r.raw.csv 2>/dev/null; done
ncu -i gpurun_out/r2_prof_gather.ncu-rep --page source --csv --print-source sass --launch-skip 0 --launch-count 1 > gpurun_out/r2_prof_gather_enc.source.csv 2>/dev/null
ncu -i gpurun_out/r2_prof_gather.ncu-rep --page source --csv --print-source sass --launch-skip 2 --launch-count 1 > gpurun_out/r2_prof_gather_dec.source.csv 2>/dev/null
rm -f gpurun_out/r2_prof_gather_split.ncu-rep gpurun_out/r2_prof_gemm.ncu-rep
ls -la gpurun_out | tail -14
