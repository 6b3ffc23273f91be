#!/bin/bash
# column-pivoted QR + preprocess on the device (one shot), then the whole suite
mkdir -p gpurun_out
( timeout 200 python -m pytest tests/test_gpu_preprocess.py -m gpu -q 2>&1 | tail -25 ) > gpurun_out/pp_pytest.log 2>&1; tail -4 gpurun_out/pp_pytest.log
( timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -6 ) > gpurun_out/pp_full.log 2>&1; tail -3 gpurun_out/pp_full.log
