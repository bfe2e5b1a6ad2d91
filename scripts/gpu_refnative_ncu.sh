#!/bin/bash
mkdir -p gpurun_out
echo "=== ref native"; timeout 900 python -m pytest tests/test_gpu_ref_native.py -m gpu -q 2>&1 | tail -30 > gpurun_out/test_ref_native.log; grep -E "passed|failed|^E  |^FAILED" gpurun_out/test_ref_native.log | cut -c1-250 | tail -25
echo "=== ncu conv (second eager step, 12 trunk convs)"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 35 -c 12 -o gpurun_out/r01a_convtrunk -f python bench.py --steps 1 --warmup 3 --no-cpu --no-graph > gpurun_out/ncu_r01a_convtrunk.log 2>&1
tail -n 1 gpurun_out/ncu_r01a_convtrunk.log | cut -c1-160
