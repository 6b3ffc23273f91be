#!/bin/bash
# stress the BASELINE bench for the intermittent chol(S) failure, then the GPU test-suite
mkdir -p gpurun_out
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-configs --no-time-to-gap"
fails=0
for rep in 1 2 3 4 5 6 7 8 9 10; do
  timeout 100 $B > gpurun_out/r_bench_$rep.json 2> gpurun_out/r_bench_$rep.err || { fails=$((fails+1)); tail -1 gpurun_out/r_bench_$rep.err | cut -c1-160; }
done
echo "== bench failures: $fails / 10" | tee gpurun_out/r_stress.log
( timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 ) > gpurun_out/r_pytest.log 2>&1
tail -4 gpurun_out/r_pytest.log
