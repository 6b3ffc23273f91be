#!/bin/bash
# the driver's N = 2 command: sphere packing (4,31) sharded by clusters + the three-point bound sharded by blocks in the same run
mkdir -p gpurun_out
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 6 --warmup 3 > gpurun_out/x_bench_2gpu.json 2> gpurun_out/x_bench_2gpu.err; echo "bench rc=$?"
tail -2 gpurun_out/x_bench_2gpu.err | cut -c1-300
python - <<PY
import json
d=json.load(open('gpurun_out/x_bench_2gpu.json')); print(round(d['ms_per_step'],3), d.get('strong_scaling')); print(d.get('sharded_configs'))
PY
