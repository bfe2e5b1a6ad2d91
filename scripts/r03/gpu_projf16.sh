#!/bin/bash
tag=${1:-q}
mkdir -p gpurun_out
o=gpurun_out/${tag}_proj_f16.txt
: > $o
timeout 600 python scripts/mobile_half_check.py 2>&1 | grep -E "trunk (half|bf16) +projections" | grep -v "trunk bf16 projections .* drn_mobilenet" | tee -a $o
timeout 600 python -m pytest tests/test_gpu_tc.py -m gpu -q -k "half_projections or deform" 2>&1 | tail -n 2 | tee -a $o
for v in 0 1; do
  TDRN_PROJ_F16=$v timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --sustain 0 --detail 2>gpurun_out/${tag}_layers_$v.txt | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('vgg320 TDRN_PROJ_F16=$v step %.4f ms deform %.4f' % (d['ms_per_step'], d['kernel_breakdown']['deform_head_tc']['ms_per_step']))" | tee -a $o
  grep -E "sample|40x40 k3\+5 project" gpurun_out/${tag}_layers_$v.txt | tee -a $o
done
