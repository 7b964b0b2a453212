#!/bin/bash
# Run on the GPU box (via gpurun): bench line, then the ncu launch list and one full capture of the
# dominant kernel on a short version of the same command. Outputs land in gpurun_out/.
set -x
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err
tail -c 3000 gpurun_out/bench_c3.json
SHORT="python bench.py --steps 2 --warmup 3 --elections-per-gpu 592 --no-cpu-baseline"
$SHORT > gpurun_out/plain_short.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $SHORT > gpurun_out/ncu_launches.log 2>&1
$SHORT > gpurun_out/plain_short2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:simulate_kernel -s 3 -c 1 -o gpurun_out/prof_simulate $SHORT > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out
