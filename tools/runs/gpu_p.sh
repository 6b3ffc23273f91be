#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-configs --no-time-to-gap"
stress() { tag=$1; shift; fails=0; for rep in 1 2 3 4 5 6; do env "$@" timeout 100 $B > /dev/null 2> gpurun_out/p_$tag.err || { fails=$((fails+1)); tail -1 gpurun_out/p_$tag.err | cut -c1-160; }; done; echo "== $tag failures: $fails / 6"; }
stress default X=1
stress nowave CLRS_TRSV_WAVEFRONT=0
stress unstaged CLRS_SCHUR_STAGED=0
stress unstaged_nowave CLRS_SCHUR_STAGED=0 CLRS_TRSV_WAVEFRONT=0
