#!/bin/bash
tag=${1:-q}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "half or dw or l2norm" 2>&1 | tail -n 25 > gpurun_out/${tag}_tests.txt; tail -n 12 gpurun_out/${tag}_tests.txt
timeout 600 python scripts/mobile_half_check.py 2>&1 | tail -n 6 | tee gpurun_out/${tag}_mobile_half.txt
timeout 900 python -m pytest tests/test_gpu_models.py tests/test_gpu_configs.py -m gpu -q -k "mobile" 2>&1 | tail -n 25 > gpurun_out/${tag}_tests_models.txt; tail -n 12 gpurun_out/${tag}_tests_models.txt
timeout 600 python bench.py --config mobilenet --steps 20 --warmup 5 --detail --no-cpu --sustain 0 > gpurun_out/${tag}_bench_mobilenet.json 2> gpurun_out/${tag}_mobilenet_layers.txt; head -n 14 gpurun_out/${tag}_mobilenet_layers.txt
python - <<PY
import json
for f in ('gpurun_out/${tag}_bench_mobilenet.json',):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, d['value'], d['ms_per_step'], d['roofline']['frac'], d.get('latency_b1'))
PY
