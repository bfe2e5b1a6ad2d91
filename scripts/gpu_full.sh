#!/bin/bash
# full GPU pass: all gpu tests, smoke, bench (+ per-layer detail, reference arm), the other BASELINE configs.
# (the ncu launch list / ncu --set full captures live in gpu_profile_all.sh)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
echo "=== tests"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -60 > gpurun_out/test_gpu.log; grep -E "passed|failed|^E   .*Error|^FAILED" gpurun_out/test_gpu.log | cut -c1-220 | tail -20
echo "=== smoke"; timeout 600 python __graft_entry__.py smoke 2>&1 | tail -12 | tee gpurun_out/smoke.log
echo "=== bench"; timeout 900 python bench.py --steps 20 --warmup 5 --detail > gpurun_out/bench.json 2> gpurun_out/bench_detail.txt; head -12 gpurun_out/bench_detail.txt; cut -c1-300 gpurun_out/bench.json
echo "=== ref"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>gpurun_out/bench_ref.err; cut -c1-300 gpurun_out/bench_ref.json
echo "=== configs"; timeout 600 python scripts/bench_configs.py > gpurun_out/bench_configs.jsonl 2> gpurun_out/bench_configs.err; cat gpurun_out/bench_configs.jsonl | cut -c1-400
echo "=== no-graph sanity"; timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu --no-graph 2>/dev/null | cut -c1-200
