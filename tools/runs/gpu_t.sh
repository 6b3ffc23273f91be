#!/bin/bash
# 1 GPU: sparse-path tests again, epilogue A/B of tc::k_gemm_tc
mkdir -p gpurun_out
( timeout 400 python -m pytest tests/test_gpu_sparse.py tests/test_gpu_parity.py -m gpu -q -x -k "sparse or gemm or tc or dsplit or swapped" 2>&1 | tail -8 ) > gpurun_out/t_pytest.log 2>&1
tail -3 gpurun_out/t_pytest.log
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-configs --no-time-to-gap"
for e in 0 1 0 1; do CLRS_TC_EPI=$e timeout 120 $B > gpurun_out/t_epi${e}.json 2> gpurun_out/t_epi${e}.err; python - <<PY
import json
d=json.load(open('gpurun_out/t_epi${e}.json')); print('epi $e', round(d['ms_per_step'],3), 'frac', round(d['roofline']['frac'],3), 'avg', round(d['roofline']['avg_launch_ms'],4), 'small', round(d['roofline']['other_gemm_classes']['tc_small']['avg_launch_ms'],4))
PY
done
