#!/bin/bash
mkdir -p gpurun_out
echo "=== tc"; timeout 600 python -m pytest tests/test_gpu_tc.py -m gpu -x -q 2>&1 | tail -30 > gpurun_out/test_gpu_tc.log; tail -6 gpurun_out/test_gpu_tc.log
echo "=== models"; timeout 900 python -m pytest tests/test_gpu_models.py -m gpu -q 2>&1 | tail -150 > gpurun_out/test_gpu_models.log; grep -E "passed|failed|^E   .*Error" gpurun_out/test_gpu_models.log | cut -c1-200 | tail -10
echo "=== bench"; timeout 900 python bench.py --steps 10 --warmup 3 --detail --no-cpu > gpurun_out/bench.json 2> gpurun_out/bench_detail.txt; head -22 gpurun_out/bench_detail.txt; python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print({k:d[k] for k in ('value','ms_per_step','e2e','roofline')}); print({k:round(v['ms_per_step'],3) for k,v in d['kernel_breakdown'].items()})"
