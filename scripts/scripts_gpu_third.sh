#!/bin/bash
mkdir -p gpurun_out
echo "=== models"; timeout 900 python -m pytest tests/test_gpu_models.py -m gpu -q 2>&1 | tail -40 > gpurun_out/test_gpu_models.log; tail -8 gpurun_out/test_gpu_models.log
echo "=== smoke"; timeout 600 python __graft_entry__.py smoke 2>&1 | tail -12 | tee gpurun_out/smoke.log
echo "=== bench"; timeout 900 python bench.py --steps 10 --warmup 3 --detail > gpurun_out/bench.json 2> gpurun_out/bench_detail.txt; cat gpurun_out/bench_detail.txt | head -70; cut -c1-400 gpurun_out/bench.json
echo "=== ncu launches"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 250 -c 130 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-graph > gpurun_out/ncu_bench.log 2>&1; tail -3 gpurun_out/launches.csv
echo "=== ncu full"; timeout 1200 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 45 -c 6 -o gpurun_out/prof_conv_tc -f python bench.py --steps 1 --warmup 3 --no-cpu --no-graph > gpurun_out/ncu_full.log 2>&1; tail -3 gpurun_out/ncu_full.log; ls -la gpurun_out/
