#!/bin/bash
# One gpurun call: A/B of library variants (graingraphnn_b200/lib/variants/*.so, GG_LIB) and env switches on the bench workload.
#   VARIANTS="default nosleep default:GG_GATHER_MERGE=0" TESTS=1 NCU=default scripts/ab_gather.sh
mkdir -p gpurun_out
if [ -n "$TESTS" ]; then timeout 1200 python -m pytest tests -m gpu -q -x --timeout 600 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log; fi
for spec in ${VARIANTS:-default}; do
  v=${spec%%:*}; envs=""; [ "$spec" != "$v" ] && envs=${spec#*:}
  tag=$(echo "$spec" | tr ':=' '__')
  lib=""; [ "$v" != default ] && lib="$PWD/graingraphnn_b200/lib/variants/$v.so"
  env GG_LIB=$lib GG_BENCH_VERBOSE=1 $envs timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-strong $BENCH_ARGS > gpurun_out/ab_$tag.json 2> gpurun_out/ab_$tag.err
  python - "$tag" <<'PY'
import json, sys
v = sys.argv[1]
try:
    d = json.loads(open(f'gpurun_out/ab_{v}.json').read().strip().splitlines()[-1])
    r = d['roofline']
    print(v, 'ms/step', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['ms_per_step'], 3), 'gather ms', round(r['ms_per_step'], 3), 'frac', round(r['frac'], 3), r['breakdown_ms'])
    print('   per gather launch us:', [round(x * 1e3) for x in r['per_call_ms'].get('gg_pgat_gather', [])])
except Exception as e:
    print(v, 'FAILED', e); print(open(f'gpurun_out/ab_{v}.err').read()[-1500:])
PY
done
if [ -n "$NCU" ]; then
  lib=""; [ "$NCU" != default ] && lib="$PWD/graingraphnn_b200/lib/variants/$NCU.so"
  GG_LIB=$lib timeout 900 ncu --set full --clock-control none --import-source on -k regex:pgat_gather -s ${NCU_SKIP:-12} -c ${NCU_COUNT:-4} -o gpurun_out/prof_gather_$NCU -f \
      python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-strong > gpurun_out/ncu_gather.log 2>&1
  tail -3 gpurun_out/ncu_gather.log
fi
