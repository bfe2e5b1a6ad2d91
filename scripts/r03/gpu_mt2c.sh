#!/bin/bash
tag=${1:-q}
o=gpurun_out/${tag}_mt2_rounds.txt
mkdir -p gpurun_out
(for v in "TDRN_MT2_ROUNDS=0" "TDRN_MT2_ROUNDS=1"; do echo "# $v"; env $v TDRN_TC_VERBOSE=1 python scripts/bench_conv.py tcb0_1 tcb0_2 tcb1_1 tcb2_1 conv4_2 conv5_1 2>&1 | sort | uniq -c | cut -c1-260; done) > $o 2>&1
cat $o
for i in 1 2; do
  for v in 0 1; do
    TDRN_MT2_ROUNDS=$v timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --sustain 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('TDRN_MT2_ROUNDS=$v step %.4f ms conv-family %.4f ms frac %.4f' % (d['ms_per_step'], d['kernel_breakdown']['conv_tc']['ms_per_step'], d['roofline']['frac']))" | tee -a $o
  done
done
