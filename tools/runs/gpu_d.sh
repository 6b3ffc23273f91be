#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -40 ) > gpurun_out/d_pytest.log 2>&1
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/d_bench.json 2> gpurun_out/d_bench.err
CLRS_TRSV_FUSED=0 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-time-to-gap > gpurun_out/d_bench_unfused.json 2> gpurun_out/d_bench_unfused.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/d_launches.csv python tools/gpu_profile_iter.py 300 3 > gpurun_out/d_launches.log 2>&1
tail -3 gpurun_out/d_pytest.log; head -c 300 gpurun_out/d_bench.json; tail -3 gpurun_out/d_bench.err
