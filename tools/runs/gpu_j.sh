#!/bin/bash
mkdir -p gpurun_out
B="--steps 5 --warmup 3 --no-cpu-baseline --no-configs --no-time-to-gap"
timeout 500 python bench.py $B --workload sphere8 > gpurun_out/j_sphere8_1.json 2> gpurun_out/j_sphere8_1.err
timeout 500 python bench.py $B --workload threepoint14 > gpurun_out/j_threepoint14_1.json 2> gpurun_out/j_threepoint14_1.err
for f in sphere8_1 threepoint14_1; do python - <<PY
import json
try:
    d=json.load(open('gpurun_out/j_$f.json')); print('$f', round(d['ms_per_step'],3), d['config']['workload'][:90], d['phase_ms'])
except Exception as e: print('$f ERR', e, open('gpurun_out/j_$f.err').read()[-600:])
PY
done
