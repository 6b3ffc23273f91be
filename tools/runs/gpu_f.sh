#!/bin/bash
mkdir -p gpurun_out
CLRS_POTRF_TIMELINE=1 python tools/gpu_chol_profile.py 64 > gpurun_out/f_potrf_tl1.log 2>&1
CLRS_POTRF_TIMELINE=2 python tools/gpu_chol_profile.py 64 > gpurun_out/f_potrf_tl2.log 2>&1
( timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -40 ) > gpurun_out/f_pytest.log 2>&1
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-time-to-gap > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err
tail -3 gpurun_out/f_pytest.log; head -c 300 gpurun_out/f_bench.json; tail -3 gpurun_out/f_bench.err; head -45 gpurun_out/f_potrf_tl1.log
