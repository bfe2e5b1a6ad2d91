#!/bin/bash
# attribution of the 1 -> 2 GPU step-time difference on ONE box: N = 1, N = 2 with and without the end-of-step gather
mkdir -p gpurun_out
o=gpurun_out/${1:-q}_scale2.txt
: > $o
show() { python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', d['n_gpus'], 'value %.0f  ms/step %.4f  e2e %.0f  per-rank %s clocks %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['per_rank_ms_per_step']['device'], d['clocks']['sm_mhz']))" | tee -a $o; }
python bench.py --steps 20 --warmup 5 --no-cpu --sustain 0 2>/dev/null | show "N=1"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu --sustain 0 2>/dev/null | show "N=2 gather"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu --sustain 0 --no-gather 2>/dev/null | show "N=2 no-gather"
python bench.py --steps 20 --warmup 5 --no-cpu --sustain 0 2>/dev/null | show "N=1 again"
