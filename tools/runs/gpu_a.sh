#!/bin/bash
# round-2 first GPU call: parity tests, bench line (graph vs eager), GEMM kernel variants
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/a_pytest.log 2>&1
python bench.py --steps 10 --warmup 3 > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err
CLRS_GRAPH=0 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-time-to-gap --no-configs > gpurun_out/a_bench_eager.json 2> gpurun_out/a_bench_eager.err
python tools/gpu_gemm_bench.py 90000x300x300 300x300x300 641x641x1294 18900x300x300 > gpurun_out/a_gemm_default.log 2>&1
CLRS_TC_GROUP=4 CLRS_TC_DSPLIT=0 python tools/gpu_gemm_bench.py 90000x300x300 300x300x300 641x641x1294 18900x300x300 > gpurun_out/a_gemm_old.log 2>&1
tail -3 gpurun_out/a_pytest.log; cat gpurun_out/a_bench.json | head -c 1500; echo; cat gpurun_out/a_gemm_default.log gpurun_out/a_gemm_old.log
