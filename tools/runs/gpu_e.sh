#!/bin/bash
# multi-GPU: parity of the sharded path and the strong-scaling bench line (run with gpurun --gpus N)
N=${1:-2}
WL=${2:-sphere}
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -x 2>&1 | tail -30 ) > gpurun_out/e_multi_pytest_$N.log 2>&1
NCCL_DEBUG=WARN timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 8 --warmup 3 --workload $WL > gpurun_out/e_bench_${WL}_$N.json 2> gpurun_out/e_bench_${WL}_$N.err
tail -5 gpurun_out/e_multi_pytest_$N.log; head -c 400 gpurun_out/e_bench_${WL}_$N.json; tail -5 gpurun_out/e_bench_${WL}_$N.err
