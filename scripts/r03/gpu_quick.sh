#!/bin/bash
# quick iteration on the stem pair / depthwise kernels: their unit tests, the issuer timing aid, both bench lines with layer detail
tag=${1:-q}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_ops.py -m gpu -x -q -k "stem or dw" 2>&1 | tail -n 15 > gpurun_out/${tag}_tests.txt; tail -n 5 gpurun_out/${tag}_tests.txt
TDRN_HALO_TIMING=1 timeout 300 python scripts/stem_pair_timing.py 2>&1 | tail -n 3 | cut -c1-400 | tee gpurun_out/${tag}_stem_timing.txt
timeout 600 python bench.py --steps 20 --warmup 5 --detail --no-cpu --sustain 0 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_layers.txt; head -n 12 gpurun_out/${tag}_layers.txt
timeout 600 python bench.py --config mobilenet --steps 20 --warmup 5 --detail --no-cpu --sustain 0 > gpurun_out/${tag}_bench_mobilenet.json 2> gpurun_out/${tag}_mobilenet_layers.txt; head -n 20 gpurun_out/${tag}_mobilenet_layers.txt
python - <<PY
import json
for f in ('gpurun_out/${tag}_bench.json','gpurun_out/${tag}_bench_mobilenet.json'):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, d['value'], d['ms_per_step'], d['roofline']['frac'], d['clocks'])
PY
