#!/bin/bash
tag=${1:-q}
mkdir -p gpurun_out
timeout 600 python scripts/mobile_half_check.py 2>&1 | tail -n 8 | tee gpurun_out/${tag}_half_check.txt
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_models.py tests/test_gpu_configs.py tests/test_gpu_properties.py -m gpu -q -x 2>&1 | tail -n 30 | cut -c1-1200 > gpurun_out/${tag}_tests.txt; tail -n 12 gpurun_out/${tag}_tests.txt
timeout 600 python bench.py --steps 20 --warmup 5 --detail --no-cpu --sustain 0 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_layers.txt; grep -E "deform|stem" gpurun_out/${tag}_layers.txt
python - <<PY
import json
d=json.loads(open('gpurun_out/${tag}_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['frac'], {k:round(v['ms_per_step'],3) for k,v in d['kernel_breakdown'].items()})
PY
