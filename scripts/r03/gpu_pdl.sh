#!/bin/bash
tag=${1:-q}
o=gpurun_out/${tag}_pdl.txt
mkdir -p gpurun_out
: > $o
TDRN_PDL=1 timeout 900 python -m pytest tests/test_gpu_models.py tests/test_gpu_configs.py tests/test_gpu_ops.py tests/test_gpu_tc.py -m gpu -q -x 2>&1 | tail -n 5 | cut -c1-300 | tee -a $o
for i in 1 2; do
for v in 0 1; do
  TDRN_PDL=$v timeout 600 python bench.py --config mobilenet --steps 20 --warmup 5 --no-cpu --sustain 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('mobilenet TDRN_PDL=$v step %.4f ms  latency_b1 p50 %.4f ms' % (d['ms_per_step'], d['latency_b1']['ms_p50']))" | tee -a $o
  TDRN_PDL=$v timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --sustain 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('vgg320 TDRN_PDL=$v step %.4f ms conv-family %.4f' % (d['ms_per_step'], d['kernel_breakdown']['conv_tc']['ms_per_step']))" | tee -a $o
done
done
