#!/bin/bash
mkdir -p gpurun_out
timeout 400 ncu --set full --import-source on --clock-control none -k regex:"k_trsv_fused|k_trsm32" --launch-skip 30 --launch-count 6 -o gpurun_out/n_trsv python tools/gpu_trsv_profile.py > gpurun_out/n_trsv.log 2>&1
tail -3 gpurun_out/n_trsv.log; ls -la gpurun_out/n_trsv*
