#!/bin/bash
# quick iteration: tc unit tests + model tests + bench with per-layer detail
mkdir -p gpurun_out
echo "=== tc"; timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_ops.py -m gpu -x -q 2>&1 | tail -n 30 > gpurun_out/test_gpu_tc.log; grep -E "passed|failed|^E  |^FAILED" gpurun_out/test_gpu_tc.log | cut -c1-250 | tail -n 15
echo "=== models"; timeout 900 python -m pytest tests/test_gpu_models.py tests/test_gpu_postprocess.py -m gpu -q 2>&1 | tail -n 60 > gpurun_out/test_gpu_models.log; grep -E "passed|failed|^E  |^FAILED" gpurun_out/test_gpu_models.log | cut -c1-250 | tail -n 15
echo "=== bench"; timeout 900 python bench.py --steps 20 --warmup 5 --detail --no-cpu > gpurun_out/bench.json 2> gpurun_out/bench_detail.txt; head -n ${1:-24} gpurun_out/bench_detail.txt; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench.json'))
print({k:d[k] for k in ('value','ms_per_step','e2e','roofline','clocks')})
print({k:round(v['ms_per_step'],3) for k,v in d['kernel_breakdown'].items()})
PY
