#!/bin/bash
mkdir -p gpurun_out
( CLRS_GRAPH=0 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "prec_512_sixteen" 2>&1 | tail -3 ) > gpurun_out/b_t512_eager.log 2>&1
( timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -40 ) > gpurun_out/b_pytest.log 2>&1
python bench.py --steps 10 --warmup 3 > gpurun_out/b_bench.json 2> gpurun_out/b_bench.err
CLRS_GRAPH=0 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-time-to-gap --no-configs > gpurun_out/b_bench_eager.json 2> gpurun_out/b_bench_eager.err
SH="19200x64x256 19200x96x256 19200x128x256 19200x144x256 19200x160x256 19200x256x256 19200x288x256 19200x320x256 19200x300x300 19200x128x1024 19200x160x1024"
python tools/gpu_gemm_bench.py $SH > gpurun_out/b_tiles_default.log 2>&1
CLRS_TC_GROUP=4 python tools/gpu_gemm_bench.py $SH > gpurun_out/b_tiles_g4.log 2>&1
CLRS_TC_GROUP=3 python tools/gpu_gemm_bench.py $SH > gpurun_out/b_tiles_g3.log 2>&1
tail -3 gpurun_out/b_t512_eager.log; tail -5 gpurun_out/b_pytest.log; head -c 600 gpurun_out/b_bench.json; tail -3 gpurun_out/b_bench.err
