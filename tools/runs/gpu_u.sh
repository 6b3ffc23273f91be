#!/bin/bash
# 2 GPUs: clusters split by blocks (three-point bound, Delsarte) against the single-GPU solve; strong scaling of config 4 (d = 10)
mkdir -p gpurun_out
CLRS_MULTI_CASES=0,3,4,5 timeout 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tests/gpu_multi.py > gpurun_out/u_multi.log 2>&1; echo "multi rc=$?"
grep -E "OK|FAIL|PARITY|Error|error" gpurun_out/u_multi.log | cut -c1-330 | head -20
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --workload threepoint --steps 6 --warmup 3 > gpurun_out/u_bench_tp_2gpu.json 2> gpurun_out/u_bench_tp_2gpu.err; echo "bench rc=$?"
tail -2 gpurun_out/u_bench_tp_2gpu.err | cut -c1-300
python - <<PY
import json
d=json.load(open('gpurun_out/u_bench_tp_2gpu.json')); print(round(d['ms_per_step'],3), d.get('strong_scaling'), {k:v for k,v in d['phase_ms'].items() if v>0.5})
PY
