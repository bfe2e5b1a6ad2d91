#!/bin/bash
tag=${1:-q}
mkdir -p gpurun_out
o=gpurun_out/${tag}_mt_major.txt
: > $o
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_ops.py -m gpu -q -x 2>&1 | tail -n 3 | cut -c1-200 | tee -a $o
TDRN_MT_MAJOR=1 timeout 900 python -m pytest tests/test_gpu_tc.py -m gpu -q -x 2>&1 | tail -n 3 | cut -c1-200 | tee -a $o
(echo "# TDRN_MT_MAJOR=0"; TDRN_MT_MAJOR=0 python scripts/pw_only.py; echo "# default (1x1 convs with >= 2 N tiles and streamed weights)"; python scripts/pw_only.py; echo "# TDRN_MT_MAJOR=1 (every layer)"; TDRN_MT_MAJOR=1 python scripts/pw_only.py) 2>&1 | tee -a $o
for v in 0 default 1; do
  if [ $v = default ]; then unset TDRN_MT_MAJOR; else export TDRN_MT_MAJOR=$v; fi
  timeout 600 python bench.py --config mobilenet --steps 20 --warmup 5 --no-cpu --sustain 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('mobilenet TDRN_MT_MAJOR=$v step %.4f ms' % (d['ms_per_step']))" | tee -a $o
  timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --sustain 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('vgg320 TDRN_MT_MAJOR=$v step %.4f ms conv-family %.4f' % (d['ms_per_step'], d['kernel_breakdown']['conv_tc']['ms_per_step']))" | tee -a $o
done
