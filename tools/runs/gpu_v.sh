#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_sparse.py tests/test_gpu_parity.py -m gpu -q -x -k "sparse or matmul_prec or gemm or tc or dsplit or swapped or triplet or bit" 2>&1 | tail -30 ) > gpurun_out/v_pytest.log 2>&1
tail -5 gpurun_out/v_pytest.log
