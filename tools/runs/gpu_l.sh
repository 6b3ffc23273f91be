#!/bin/bash
mkdir -p gpurun_out
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file gpurun_out/l_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-time-to-gap --no-configs > gpurun_out/l_bench_under_ncu.json 2> gpurun_out/l_bench_under_ncu.err
timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_gemm_tc --launch-count 6 -o gpurun_out/l_tc python tools/gpu_profile_iter.py 300 1 > gpurun_out/l_ncu_tc.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:"k_tc_recombine|k_split_tc|k_vec_exp|k_gemm_dp4a" --launch-count 14 -o gpurun_out/l_hbm python tools/gpu_profile_iter.py 300 1 > gpurun_out/l_ncu_hbm.log 2>&1
ls -la gpurun_out/l_*; tail -2 gpurun_out/l_ncu_tc.log
