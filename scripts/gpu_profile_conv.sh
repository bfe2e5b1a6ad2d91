#!/bin/bash
# Partial profile refresh (kernels that changed since the last full gpu_profile_all.sh set): launch list of one eager step,
# ncu --set full of the conv family (35 launches, deformable heads on their im2col path so that the projection GEMMs stay
# out of the family's traffic mean) and of the fused stem pair with source-level sampling.
# usage: gpu_profile_conv.sh <tag>
mkdir -p gpurun_out /tmp/prof
tag=${1:-r01}
B="python bench.py --steps 1 --warmup 3 --no-cpu --no-graph"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_launches.csv $B > /tmp/prof/launches.log 2>&1
python scripts/launch_summary.py gpurun_out/${tag}_launches.csv 45 > gpurun_out/${tag}_launches.txt
run() { timeout 1500 ncu --set full --clock-control none $5 -k regex:$2 -s $3 -c $4 -o /tmp/prof/${tag}_$1 -f $B > /tmp/prof/ncu_$1.log 2>&1; tail -n 1 /tmp/prof/ncu_$1.log | cut -c1-120; python scripts/ncu_table.py /tmp/prof/${tag}_$1.ncu-rep > gpurun_out/${tag}_ncu_$1.txt; }
TDRN_DEFORM_PATH=im2col run conv 'conv_(tc|halo|stem_pair)' 35 35 ""
python scripts/ncu_traffic.py /tmp/prof/${tag}_conv.ncu-rep gpurun_out/${tag}_conv_traffic.json
run stempair 'conv_stem_pair_kernel' 1 1 "--import-source on"
python scripts/ncu_summary.py /tmp/prof/${tag}_stempair.ncu-rep >> gpurun_out/${tag}_ncu_stempair.txt
cp /tmp/prof/${tag}_stempair.ncu-rep gpurun_out/ 2>/dev/null
head -12 gpurun_out/${tag}_launches.txt
