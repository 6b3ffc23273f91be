#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -40 ) > gpurun_out/c_pytest.log 2>&1
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c_bench.json 2> gpurun_out/c_bench.err
CLRS_CHOL_PANEL=32 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-time-to-gap > gpurun_out/c_bench_panel32.json 2> gpurun_out/c_bench_panel32.err
ncu --set full --import-source on --clock-control none -k regex:"k_potrf_diag|k_trsv_block|k_trsm32" -c 12 -o gpurun_out/c_panel python tools/gpu_chol_profile.py 300 > gpurun_out/c_ncu_panel.log 2>&1
tail -3 gpurun_out/c_pytest.log; head -c 300 gpurun_out/c_bench.json; tail -3 gpurun_out/c_bench.err; tail -3 gpurun_out/c_ncu_panel.log
