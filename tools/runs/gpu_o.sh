#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-configs --no-time-to-gap"
for rep in 1 2; do
timeout 200 $B > gpurun_out/o_wave_$rep.json 2> gpurun_out/o_wave_$rep.err; echo "wave $rep rc=$?"; tail -1 gpurun_out/o_wave_$rep.err | cut -c1-200
CLRS_TRSV_WAVEFRONT=0 timeout 200 $B > gpurun_out/o_nowave_$rep.json 2> gpurun_out/o_nowave_$rep.err; echo "nowave $rep rc=$?"; tail -1 gpurun_out/o_nowave_$rep.err | cut -c1-200
done
for f in wave_1 wave_2 nowave_1 nowave_2; do python - <<PY
import json
try:
    d=json.load(open('gpurun_out/o_$f.json')); print('$f', round(d['ms_per_step'],3), round(d['e2e']['value'],2), {k:v for k,v in d['phase_ms'].items() if k in ('solve','cholS','Xinv','schur')})
except Exception as e: print('$f ERR')
PY
done
