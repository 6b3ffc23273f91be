#!/bin/bash
# k_trsv_fused with per-row flags instead of per-column barriers: test-suite + bench (phase timer `solve` of every config)
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 ) > gpurun_out/y_pytest.log 2>&1; tail -3 gpurun_out/y_pytest.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/y_bench.json 2> gpurun_out/y_bench.err; echo "bench rc=$?"; tail -1 gpurun_out/y_bench.err | cut -c1-200
python - <<'PY'
import json
d=json.load(open('gpurun_out/y_bench.json'))
print(round(d['ms_per_step'],3), round(d['value'],2), 'e2e', round(d['e2e']['value'],2), 'frac', round(d['roofline']['frac'],3), 'solve', d['phase_ms']['solve'], d.get('time_to_gap_1e-30',{}).get('seconds'))
for k,v in d.get('configs',{}).items(): print(k, round(v.get('ms_per_step',0),3), 'solve', v.get('phase_ms',{}).get('solve'))
PY
