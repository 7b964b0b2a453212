#!/bin/bash
# ncu --set full of the stand-alone Gamma and binomial kernels (debug hooks) -> gpurun_out/r02_ncu_<stage>_stage.txt
for stage in gamma binomial; do
  python tools/variate_probe.py $stage || exit 1
  ncu --set full --clock-control none --import-source on -k regex:${stage}_kernel -s 1 -c 3 -f -o gpurun_out/prof_$stage python tools/variate_probe.py $stage > gpurun_out/ncu_$stage.log 2>&1
  for id in 0 1 2; do
    ncu -i gpurun_out/prof_$stage.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); hdr,units=rows[0],rows[1]
for vals in rows[2:]:
    g=lambda k: next((vals[i]+' '+units[i] for i,h in enumerate(hdr) if h==k),'n/a')
    print('launch', vals[0], g('Kernel Name')[:50])
    for k in ('gpu__time_duration.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','smsp__thread_inst_executed_per_inst_executed.ratio','sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','sm__warps_active.avg.pct_of_peak_sustained_active','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','launch__registers_per_thread'):
        print('   %-72s %s' % (k, g(k)))
" > gpurun_out/r02_ncu_${stage}_stage.txt
    break
  done
  rm -f gpurun_out/prof_$stage.ncu-rep
  cat gpurun_out/r02_ncu_${stage}_stage.txt
done
