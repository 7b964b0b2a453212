#!/bin/bash
# ncu --set full capture of one launch of a kernel inside a tools/heavy_probe.py run; summary into gpurun_out/.
# usage: tools/profile_case.sh TAG KERNEL_REGEX SKIP -- heavy_probe args...
TAG=$1; KERNEL=$2; SKIP=$3; shift 4
python tools/heavy_probe.py "$@" > gpurun_out/${TAG}_plain.log 2>&1 || { tail gpurun_out/${TAG}_plain.log; exit 1; }
tail -2 gpurun_out/${TAG}_plain.log
ncu --set full --clock-control none --import-source on -k regex:$KERNEL -s $SKIP -c 1 -f -o gpurun_out/prof_$TAG \
  python tools/heavy_probe.py "$@" > gpurun_out/${TAG}_ncu.log 2>&1
tail -2 gpurun_out/${TAG}_ncu.log
python tools/ncu_summary.py gpurun_out/prof_$TAG.ncu-rep 45 > gpurun_out/${TAG}_summary.txt 2>&1
rm -f gpurun_out/prof_$TAG.ncu-rep
head -70 gpurun_out/${TAG}_summary.txt
