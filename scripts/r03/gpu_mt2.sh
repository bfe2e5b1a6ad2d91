#!/bin/bash
tag=${1:-q}
mkdir -p gpurun_out
o=gpurun_out/${tag}_mt2_experiment.txt
(echo "# conv5_1 (512->512 @20x20 b32, 100 M tiles x 2 N tiles): default (MT2 from 70 units per 100 SMs)"; python scripts/bench_conv.py conv5_1 conv4_2
 echo "# TDRN_MT2_MIN=60 (pairs two M tiles per weight box on conv5: 100 units)"; TDRN_MT2_MIN=60 python scripts/bench_conv.py conv5_1 conv4_2
 echo "# TDRN_CLUSTER=1 (weight boxes multicast over CTA pairs)"; TDRN_CLUSTER=1 python scripts/bench_conv.py conv5_1 conv4_2) > $o 2>&1
cat $o
for i in 1 2; do
  for v in 70 60; do
    TDRN_MT2_MIN=$v timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --sustain 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('TDRN_MT2_MIN=$v step %.4f ms conv-family %.4f ms frac %.4f' % (d['ms_per_step'], d['kernel_breakdown']['conv_tc']['ms_per_step'], d['roofline']['frac']))" | tee -a $o
  done
done
