#!/bin/bash
# the driver's N = 4 command (sphere packing by clusters + three-point bound by blocks in one run)
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 4 --steps 4 --warmup 3 > gpurun_out/x4_bench_4gpu.json 2> gpurun_out/x4_bench_4gpu.err; echo "bench rc=$?"
tail -2 gpurun_out/x4_bench_4gpu.err | cut -c1-300
python - <<PY
import json
d=json.load(open('gpurun_out/x4_bench_4gpu.json')); print(round(d['ms_per_step'],3), d.get('strong_scaling')); t=d.get('sharded_configs',{}).get('threepoint',{}); print({k:t.get(k) for k in ('ms_per_step','ms_per_step_1gpu','speedup','ranks_holding_blocks','error')})
PY
