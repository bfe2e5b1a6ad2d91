#!/bin/bash
tag=${1:-q}
mkdir -p gpurun_out
o=gpurun_out/${tag}_epilogue.txt
: > $o
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_ops.py -m gpu -q -x 2>&1 | tail -n 3 | cut -c1-200 | tee -a $o
python scripts/pw_only.py 2>&1 | tee -a $o
for i in 1 2; do
  timeout 600 python bench.py --config mobilenet --steps 20 --warmup 5 --no-cpu --sustain 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('mobilenet step %.4f ms latency_b1 %.4f' % (d['ms_per_step'], d['latency_b1']['ms_p50']))" | tee -a $o
  timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --sustain 0 --detail 2>gpurun_out/${tag}_layers.txt | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('vgg320 step %.4f ms conv-family %.4f frac %.4f deform %.4f' % (d['ms_per_step'], d['kernel_breakdown']['conv_tc']['ms_per_step'], d['roofline']['frac'], d['kernel_breakdown']['deform_head_tc']['ms_per_step']))" | tee -a $o
done
grep project gpurun_out/${tag}_layers.txt | tee -a $o
