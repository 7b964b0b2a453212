#!/bin/bash
# One ncu --set full capture of lane_kernel on the bench's C3 step (run via gpurun; one profiler per call).
cd "$(dirname "$0")/.."
SHORT="python bench.py --steps 2 --warmup 3 --elections-per-gpu ${EPG:-100000} --no-cpu-baseline --single-mode"
timeout 300 $SHORT > gpurun_out/plain_short.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_short.log; exit 1; }
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lane_kernel -s 3 -c 1 -f -o gpurun_out/prof_${TAG:-lane} $SHORT > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out/*.ncu-rep
