#!/bin/bash
# Run on the GPU box (via gpurun): bench line, then the ncu launch list and one full capture of the
# dominant kernel on a short version of the same command. Outputs land in gpurun_out/.
# usage: tools/profile_c3.sh [tag] [extra bench args...]
TAG=${1:-c3}; shift
set -x
python bench.py --steps 3 --warmup 3 "$@" > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
tail -c 2500 gpurun_out/bench_$TAG.json
SHORT="python bench.py --steps 2 --warmup 3 --elections-per-gpu ${EPG:-592} --no-cpu-baseline --single-mode $*"
$SHORT > gpurun_out/plain_short.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv $SHORT > gpurun_out/ncu_launches.log 2>&1
$SHORT > gpurun_out/plain_short2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:${KERNEL:-simulate_kernel} -s 3 -c 1 -o gpurun_out/prof_$TAG $SHORT > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
