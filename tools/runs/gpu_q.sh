#!/bin/bash
mkdir -p gpurun_out
timeout 500 compute-sanitizer --tool racecheck --racecheck-report all python tools/gpu_racecheck.py > gpurun_out/q_racecheck.log 2>&1
grep -c "Race reported\|hazard" gpurun_out/q_racecheck.log; grep -A3 "hazard" gpurun_out/q_racecheck.log | head -40; tail -5 gpurun_out/q_racecheck.log
