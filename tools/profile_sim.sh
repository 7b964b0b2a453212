#!/bin/bash
# ncu evidence for the materialising path as shipped (three phase kernels per wave): launch list of one
# tools/phase_probe.py run, then one --set full capture of each phase kernel and of the fused simulate_kernel.
# usage (through gpurun, one GPU): tools/profile_sim.sh TAG
TAG=${1:-r02s}
python tools/phase_probe.py C3 3552 0 > gpurun_out/${TAG}_plain.json 2> gpurun_out/${TAG}_plain.err || { tail gpurun_out/${TAG}_plain.err; exit 1; }
cat gpurun_out/${TAG}_plain.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${TAG}_launches_sim.csv \
  python tools/phase_probe.py C3 3552 0 > gpurun_out/${TAG}_ncu_launches.log 2>&1
for k in sim_big sim_walk sim_irv; do
  ncu --set full --clock-control none --import-source on -k regex:$k -c 1 -f -o gpurun_out/prof_${TAG}_$k \
    python tools/phase_probe.py C3 1776 0 > gpurun_out/${TAG}_ncu_$k.log 2>&1
  python tools/ncu_summary.py gpurun_out/prof_${TAG}_$k.ncu-rep 40 > gpurun_out/${TAG}_${k}_kernel_c3.txt 2>&1
  rm -f gpurun_out/prof_${TAG}_$k.ncu-rep
  head -12 gpurun_out/${TAG}_${k}_kernel_c3.txt
done
ncu --set full --clock-control none --import-source on -k regex:simulate_kernel -s 1 -c 1 -f -o gpurun_out/prof_${TAG}_fused \
  python tools/phase_probe.py C3 592 0 sim_phased=0 > gpurun_out/${TAG}_ncu_fused.log 2>&1
python tools/ncu_summary.py gpurun_out/prof_${TAG}_fused.ncu-rep 40 > gpurun_out/${TAG}_simulate_kernel_fused_c3.txt 2>&1
rm -f gpurun_out/prof_${TAG}_fused.ncu-rep
head -12 gpurun_out/${TAG}_simulate_kernel_fused_c3.txt
