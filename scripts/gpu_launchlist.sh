#!/bin/bash
# ncu launch list (device time per launch) of the LAST eager step of a short bench run
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-graph > gpurun_out/ncu_bench.log 2>&1
wc -l gpurun_out/launches.csv
