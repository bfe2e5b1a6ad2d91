#!/bin/bash
tag=${1:-q}
mkdir -p gpurun_out
o=gpurun_out/${tag}_halo_epilogue.txt
: > $o
timeout 900 python -m pytest tests/test_gpu_tc.py -m gpu -q -x 2>&1 | tail -n 3 | cut -c1-200 | tee -a $o
python scripts/bench_conv.py conv1_2 conv2_1 conv2_2 conv3_1 conv3_2 2>&1 | tee -a $o
TDRN_HALO_TIMING=1 python scripts/bench_conv.py conv2_1 2>&1 | grep -i "halo timing" | tail -n 1 | cut -c1-300 | tee -a $o
for i in 1 2; do
  timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --sustain 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('vgg320 step %.4f ms conv-family %.4f frac %.4f' % (d['ms_per_step'], d['kernel_breakdown']['conv_tc']['ms_per_step'], d['roofline']['frac']))" | tee -a $o
done
