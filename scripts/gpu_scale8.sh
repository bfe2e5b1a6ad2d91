#!/bin/bash
# one 8-GPU box, back to back: N = 1 and N = 8 of the default workload (the driver computes the scaling efficiency from its own runs)
mkdir -p gpurun_out
o=gpurun_out/${1:-q}_scale8.txt
: > $o
show() { python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', d['n_gpus'], 'value %.0f  ms/step %.4f  e2e %.0f  per-rank %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], [round(v,4) for v in d['per_rank_ms_per_step']['device']]))" | tee -a $o; }
python bench.py --steps 20 --warmup 5 --no-cpu --sustain 0 2>/dev/null | show "N=1"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu --sustain 0 2>/dev/null > gpurun_out/${1:-q}_bench_8gpu.json; cat gpurun_out/${1:-q}_bench_8gpu.json | show "N=8"
