#!/bin/bash
mkdir -p gpurun_out
echo "=== tc"; timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_ops.py -m gpu -x -q 2>&1 | tail -30 > gpurun_out/test_gpu_tc.log; tail -12 gpurun_out/test_gpu_tc.log
echo "=== models"; timeout 900 python -m pytest tests/test_gpu_models.py -m gpu -q 2>&1 | tail -150 > gpurun_out/test_gpu_models.log; grep -E "passed|failed|Error:|^E  " gpurun_out/test_gpu_models.log | cut -c1-200 | tail -30
echo "=== detect micro"; timeout 600 python scripts/bench_detect.py 2>&1 | tail -12 | tee gpurun_out/bench_detect.log
echo "=== bench"; timeout 900 python bench.py --steps 10 --warmup 3 --detail --no-cpu > gpurun_out/bench.json 2> gpurun_out/bench_detail.txt; head -45 gpurun_out/bench_detail.txt; python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print({k:d[k] for k in ('value','ms_per_step','e2e','roofline')}); print(d['kernel_breakdown'])"
