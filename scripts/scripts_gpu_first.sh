#!/bin/bash
# first GPU bring-up: SIMT ops + postprocess, then tcgen05 unit tests, then whole models
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
for f in test_gpu_ops test_gpu_postprocess test_gpu_tc test_gpu_models; do
  echo "=== $f" 
  timeout 600 python -m pytest tests/$f.py -m gpu -x -q 2>&1 | tail -40 > gpurun_out/$f.log
  tail -15 gpurun_out/$f.log
done
