#!/bin/bash
# ncu --set full of representative launches of one eager step (never a bench number).
# usage: gpu_ncu_full.sh <tag>
mkdir -p gpurun_out
tag=${1:-r01}
run() { # name regex skip count
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c $4 -o gpurun_out/${tag}_$1 -f \
     python bench.py --steps 1 --warmup 3 --no-cpu --no-graph > gpurun_out/ncu_${tag}_$1.log 2>&1
  tail -1 gpurun_out/ncu_${tag}_$1.log | cut -c1-160
}
# bench.py runs 3 eager warm-ups + 1 count + capture + timed ... : skip the first steps' launches
run conv64  'conv_tc_kernel<64>'  10 2
run conv128 'conv_tc_kernel<128>' 4 2
run conv256 'conv_tc_kernel<256>' 56 6
run deform  'deform_head_kernel'  8 4
run nms     'nms_segment_kernel'  2 1
run first   'conv_first_kernel'   2 1
ls -la gpurun_out/*.ncu-rep
