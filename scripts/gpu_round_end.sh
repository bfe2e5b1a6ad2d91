#!/bin/bash
# Round-end evidence in one gpurun call (one GPU, ~12 minutes): the whole GPU test-suite, smoke(), the ncu launch list + conv-family
# traffic capture of THIS build (profiles/conv_traffic.json carries the hash of the kernel sources), and the bench lines of all
# four BASELINE.json configurations with their cpu_baseline legs.   usage: gpu_round_end.sh <tag>
mkdir -p gpurun_out
tag=${1:-r02}
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -n 15 > gpurun_out/${tag}_pytest_gpu.txt
timeout 600 python __graft_entry__.py smoke > gpurun_out/${tag}_smoke.txt 2>&1
bash scripts/gpu_profile_conv.sh ${tag} > /dev/null 2>&1
python bench.py --steps 20 --warmup 5 --detail > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench_layers.txt
python bench.py --steps 20 --warmup 5 --impl reference > gpurun_out/${tag}_bench_reference.json 2> /dev/null
for c in coco512 mobilenet tdrn; do python bench.py --config $c --steps 20 --warmup 5 > gpurun_out/${tag}_bench_$c.json 2> /dev/null; done
python bench.py --precision fp32 --steps 10 --warmup 3 --no-cpu --sustain 0 --detail > gpurun_out/${tag}_bench_fp32.json 2> gpurun_out/${tag}_bench_fp32_layers.txt
python bench.py --config mobilenet --steps 10 --warmup 3 --no-cpu --sustain 0 --detail > /dev/null 2> gpurun_out/${tag}_bench_mobilenet_layers.txt
python scripts/dwpw_timing.py > gpurun_out/${tag}_dwpw_timing.txt 2>&1
python scripts/bench_detect.py > gpurun_out/${tag}_bench_detect.txt 2>&1
tail -n 4 gpurun_out/${tag}_pytest_gpu.txt; tail -n 6 gpurun_out/${tag}_smoke.txt; head -n 8 gpurun_out/${tag}_launches.txt
python - <<PY
import json
for c in ("", "_coco512", "_mobilenet", "_tdrn", "_fp32"):
    try:
        d = json.load(open("gpurun_out/${tag}_bench%s.json" % c))
        print(c, round(d["value"]), round(d["ms_per_step"], 3), round(d["e2e"]["value"]), d["sustained"]["value"] if d["sustained"] else None,
              d["roofline"]["frac"], d["roofline"]["traffic"], d["cpu_baseline"]["value"] if d["cpu_baseline"] else None)
    except Exception as e:
        print(c, "ERR", e)
PY
