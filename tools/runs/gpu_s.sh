#!/bin/bash
# round 2, batch s: sparse Schur path + SDPA triplets + pipelined GEMM epilogue
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -25 ) > gpurun_out/s_pytest.log 2>&1
tail -6 gpurun_out/s_pytest.log
timeout 500 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/s_bench.json 2> gpurun_out/s_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/s_bench.err | cut -c1-300
python - <<'PY'
import json
d=json.load(open('gpurun_out/s_bench.json'))
print(round(d['ms_per_step'],3), round(d['value'],2), 'e2e', round(d['e2e']['value'],2), 'frac', round(d['roofline']['frac'],3), 'avg', round(d['roofline']['avg_launch_ms'],4), d.get('time_to_gap_1e-30'))
print({k:v for k,v in d['phase_ms'].items() if v>0.2})
for k,v in d.get('configs',{}).items(): print(k, {a:(round(b,3) if isinstance(b,float) else b) for a,b in v.items() if a not in('phase_ms','workload','note')})
PY
