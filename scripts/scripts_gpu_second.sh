#!/bin/bash
mkdir -p gpurun_out
echo "=== models"; timeout 900 python -m pytest tests/test_gpu_models.py -m gpu -q 2>&1 | tail -40 > gpurun_out/test_gpu_models.log; tail -25 gpurun_out/test_gpu_models.log
echo "=== smoke"; timeout 600 python __graft_entry__.py smoke 2>&1 | tail -20 | tee gpurun_out/smoke.log
echo "=== bench"; timeout 900 python bench.py --steps 10 --warmup 3 2>&1 | tail -5 | tee gpurun_out/bench.log
