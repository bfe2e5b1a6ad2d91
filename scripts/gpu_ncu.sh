#!/bin/bash
# ncu --set full captures of selected kernels of one eager bench step (kernel regex, skip, count as args)
mkdir -p gpurun_out
name=$1; regex=$2; skip=$3; count=$4
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:$regex -s $skip -c $count -o gpurun_out/prof_$name -f python bench.py --steps 1 --warmup 3 --no-cpu --no-graph > gpurun_out/ncu_$name.log 2>&1
tail -2 gpurun_out/ncu_$name.log | cut -c1-200
