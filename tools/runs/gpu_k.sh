#!/bin/bash
# 8-GPU run: 4-rank parity of the sharded path, strong scaling of sphere packing (8,40) and (4,31)
mkdir -p gpurun_out
( timeout 290 python -m pytest tests/test_gpu_multi.py -m gpu -q -k "4-20" 2>&1 | tail -15 ) > gpurun_out/k_multi_pytest_4.log 2>&1
run() { N=$1; WL=$2; T=$3; NCCL_DEBUG=WARN timeout $T python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$N bench.py --gpus $N --steps 5 --warmup 3 --workload $WL > gpurun_out/k_bench_${WL}_$N.json 2> gpurun_out/k_bench_${WL}_$N.err; }
run 8 sphere8 330
run 8 sphere 150
tail -3 gpurun_out/k_multi_pytest_4.log
for f in sphere8_8 sphere_8; do python - <<PY
import json
try:
    d=json.load(open('gpurun_out/k_bench_$f.json')); print('$f', round(d['ms_per_step'],3), d.get('strong_scaling'), d['phase_ms'])
except Exception as e: print('$f ERR', e, open('gpurun_out/k_bench_$f.err').read()[-800:])
PY
done
