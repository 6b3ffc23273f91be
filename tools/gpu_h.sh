#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -40 ) > gpurun_out/h_pytest.log 2>&1
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/h_bench.json 2> gpurun_out/h_bench.err
CLRS_WOPS_BENCH=1 timeout 100 python tools/gpu_wops.py 2>&1 | tail -6 > gpurun_out/h_wops.log
tail -3 gpurun_out/h_pytest.log; head -c 300 gpurun_out/h_bench.json; tail -3 gpurun_out/h_bench.err; cat gpurun_out/h_wops.log
