#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -40 ) > gpurun_out/h_pytest.log 2>&1
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/h_bench.json 2> gpurun_out/h_bench.err
CLRS_SCHUR_STAGED=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-configs --no-time-to-gap > gpurun_out/h_bench_unstaged.json 2> gpurun_out/h_bench_unstaged.err
CLRS_GRAPH=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-configs --no-time-to-gap > gpurun_out/h_bench_eager.json 2> gpurun_out/h_bench_eager.err
tail -3 gpurun_out/h_pytest.log; head -c 300 gpurun_out/h_bench.json; tail -3 gpurun_out/h_bench.err
