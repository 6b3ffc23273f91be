#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -40 ) > gpurun_out/i_pytest.log 2>&1
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-configs --no-time-to-gap"
timeout 200 $B > gpurun_out/i_persist.json 2> gpurun_out/i_persist.err
CLRS_TC_PERSISTENT=0 timeout 200 $B > gpurun_out/i_nopersist.json 2> gpurun_out/i_nopersist.err
python tools/gpu_gemm_bench.py 90000x300x300 18900x300x300 > gpurun_out/i_gemm_persist.log 2>&1
CLRS_TC_PERSISTENT=0 python tools/gpu_gemm_bench.py 90000x300x300 18900x300x300 > gpurun_out/i_gemm_nopersist.log 2>&1
tail -3 gpurun_out/i_pytest.log
for f in persist nopersist; do python - <<PY
import json
try:
    d=json.load(open('gpurun_out/i_$f.json')); print('$f', round(d['ms_per_step'],3), round(d['roofline']['frac'],3), {k:v for k,v in d['phase_ms'].items() if k in ('schur','Xinv','cholS','decomp')})
except Exception as e: print('$f ERR', e)
PY
done
cat gpurun_out/i_gemm_persist.log gpurun_out/i_gemm_nopersist.log
