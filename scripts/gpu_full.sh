#!/bin/bash
# full GPU pass: all gpu tests, smoke, bench (+ per-layer detail), ncu launch list of one eager step
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
echo "=== tests"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -60 > gpurun_out/test_gpu.log; grep -E "passed|failed|^E   .*Error|^FAILED" gpurun_out/test_gpu.log | cut -c1-220 | tail -20
echo "=== smoke"; timeout 600 python __graft_entry__.py smoke 2>&1 | tail -12 | tee gpurun_out/smoke.log
echo "=== bench"; timeout 900 python bench.py --steps 20 --warmup 5 --detail > gpurun_out/bench.json 2> gpurun_out/bench_detail.txt; head -60 gpurun_out/bench_detail.txt; cut -c1-600 gpurun_out/bench.json
echo "=== ref"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>gpurun_out/bench_ref.err; cut -c1-300 gpurun_out/bench_ref.json
echo "=== ncu launches"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-graph > gpurun_out/ncu_bench.log 2>&1; wc -l gpurun_out/launches.csv
