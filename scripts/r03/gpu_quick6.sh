#!/bin/bash
tag=${1:-q}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py -m gpu -x -q -k "stem" 2>&1 | tail -n 5 | cut -c1-300 | tee gpurun_out/${tag}_tests.txt
TDRN_HALO_TIMING=1 timeout 300 python scripts/stem_pair_timing.py 2>&1 | tail -n 2 | cut -c1-400 | tee gpurun_out/${tag}_stem_timing.txt
for i in 1 2; do
timeout 600 python bench.py --steps 20 --warmup 5 --detail --no-cpu --sustain 0 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_layers.txt; grep -E "stem" gpurun_out/${tag}_layers.txt
python - <<PY
import json
d=json.loads(open('gpurun_out/${tag}_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['frac'], {k:round(v['ms_per_step'],3) for k,v in d['kernel_breakdown'].items()})
PY
done
