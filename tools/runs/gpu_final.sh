#!/bin/bash
# end of round 2: test-suite, the default bench line, ncu launch list of the bench command, full captures of the top kernels
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 ) > gpurun_out/z_pytest.log 2>&1; tail -3 gpurun_out/z_pytest.log
timeout 600 python bench.py > gpurun_out/z_bench.json 2> gpurun_out/z_bench.err; echo "bench rc=$?"; tail -1 gpurun_out/z_bench.err | cut -c1-200
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/z_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-time-to-gap --no-configs > gpurun_out/z_bench_under_ncu.json 2> gpurun_out/z_bench_under_ncu.err; echo "ncu list rc=$?"
true
timeout 300 ncu --set full --clock-control none -k regex:"k_trsm32|k_tc_recombine|k_split_tc_tri|k_trsv_fused" --launch-count 10 -o gpurun_out/z_chain python tools/gpu_profile_iter.py 300 1 > gpurun_out/z_ncu_chain.log 2>&1; echo "ncu chain rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/z_bench.json'))
print(round(d['ms_per_step'],3), round(d['value'],2), 'e2e', round(d['e2e']['value'],2), 'frac', round(d['roofline']['frac'],3), d.get('time_to_gap_1e-30'), d.get('cpu_baseline'))
for k,v in d.get('configs',{}).items(): print(k, round(v.get('ms_per_step',0),3))
PY
ls -la gpurun_out/z_*
