#!/bin/bash
# last sanity of round 2: smoke(), the new Cohn-Elkies device test, a handful of parity tests on the rebuilt library
mkdir -p gpurun_out
timeout 200 python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/zz_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/zz_smoke.log | cut -c1-250
( timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sparse.py -m gpu -q -k "cohn or three_cycle or gemm_vs_oracle or cholesky_vs or delsarte_e8 or sparse_schur_solves or golden" 2>&1 | tail -5 ) > gpurun_out/zz_pytest.log 2>&1; tail -3 gpurun_out/zz_pytest.log
