#!/bin/bash
mkdir -p gpurun_out
CLRS_POTRF_TIMELINE=1 timeout 100 python tools/gpu_chol_profile.py 64 2>&1 | tail -3 > gpurun_out/m_potrf_tl.log
( timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -40 ) > gpurun_out/m_pytest.log 2>&1
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/m_bench.json 2> gpurun_out/m_bench.err
tail -4 gpurun_out/m_pytest.log; head -c 300 gpurun_out/m_bench.json; tail -3 gpurun_out/m_bench.err; cat gpurun_out/m_potrf_tl.log
