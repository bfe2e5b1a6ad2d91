#!/bin/bash
# ncu --set full of the project-then-sample pair at pyramid level 0 (b32, 40x40): one launch of each kernel
# usage: gpu_deform_prof.sh <tag>
mkdir -p gpurun_out /tmp/prof
tag=${1:-r01d}
export LEVELS=40 CHUNKS=100000
for k in deform_sample_kernel conv_tc_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 6 -c 1 -o /tmp/prof/${tag}_$k -f python scripts/bench_deform.py > /tmp/prof/ncu_$k.log 2>&1
  tail -n 1 /tmp/prof/ncu_$k.log | cut -c1-120
  python scripts/ncu_summary.py /tmp/prof/${tag}_$k.ncu-rep > gpurun_out/${tag}_ncu_$k.txt
  cp /tmp/prof/${tag}_$k.ncu-rep gpurun_out/
done
