#!/bin/bash
mkdir -p gpurun_out
echo "=== post"; timeout 600 python -m pytest tests/test_gpu_postprocess.py -m gpu -x -q 2>&1 | tail -30 > gpurun_out/test_gpu_post.log; tail -12 gpurun_out/test_gpu_post.log
echo "=== detect micro"; timeout 600 python scripts/bench_detect.py 2>&1 | tail -12 | tee gpurun_out/bench_detect.log
echo "=== smoke"; timeout 600 python __graft_entry__.py smoke 2>&1 | tail -4
