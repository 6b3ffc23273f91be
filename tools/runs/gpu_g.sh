#!/bin/bash
mkdir -p gpurun_out
run() { tag=$1; shift; ( env "$@" timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 tools/gpu_multi_quick.py 6 2>&1 | grep -v "^\*\|OMP_NUM" | tail -20 ) > gpurun_out/g_$tag.log 2>&1; echo "== $tag rc=$?"; tail -4 gpurun_out/g_$tag.log; }
run eager_nolanes_unfused CLRS_GRAPH=0 CLRS_LANES_MIN=100000000 CLRS_TRSV_FUSED=0
run graph_only CLRS_LANES_MIN=100000000 CLRS_TRSV_FUSED=0
run lanes_only CLRS_GRAPH=0 CLRS_TRSV_FUSED=0
run fused_only CLRS_GRAPH=0 CLRS_LANES_MIN=100000000
run all_default X=1
CLRS_WOPS_BENCH=1 python tools/gpu_wops.py 2>&1 | tail -6 > gpurun_out/g_wops.log; cat gpurun_out/g_wops.log
