#!/bin/bash
# ncu launch list (gpu__time_duration per launch) of the short bench command; one profiler per gpurun call.
cd "$(dirname "$0")/.."
SHORT="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --single-mode"
timeout 200 $SHORT > gpurun_out/plain_short.log 2>&1 || { echo "plain run failed"; exit 1; }
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_${TAG:-lane}.csv $SHORT > gpurun_out/ncu_launches.log 2>&1
grep -c lane_kernel gpurun_out/launches_${TAG:-lane}.csv
