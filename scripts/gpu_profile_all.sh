#!/bin/bash
# The profile set committed under profiles/ for a round: launch list + ncu --set full of the conv family (one whole
# eager step), the deformable heads (projection GEMM + sampler), the stem and the post-processing kernels.  Summaries are
# produced ON the box (the .ncu-rep files of 35 launches exceed what gpurun copies back); only text / json and small
# reports return.
# usage: gpu_profile_all.sh <tag>
mkdir -p gpurun_out /tmp/prof
tag=${1:-r01}
B="python bench.py --steps 1 --warmup 3 --no-cpu --no-graph"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_launches.csv $B > /tmp/prof/launches.log 2>&1
python scripts/launch_summary.py gpurun_out/${tag}_launches.csv 61 > gpurun_out/${tag}_launches.txt
run() { timeout 1500 ncu --set full --clock-control none $5 -k regex:$2 -s $3 -c $4 -o /tmp/prof/${tag}_$1 -f $B > /tmp/prof/ncu_$1.log 2>&1; tail -n 1 /tmp/prof/ncu_$1.log | cut -c1-120; python scripts/ncu_table.py /tmp/prof/${tag}_$1.ncu-rep > gpurun_out/${tag}_ncu_$1.txt; }
# conv family = the 35 conv launches of the step: capture it with the deformable heads on their im2col path so that the
# per-tap projection GEMMs (same kernel name, reported with the deformable head) stay out of the family's DRAM-traffic mean
TDRN_DEFORM_PATH=im2col run conv   'conv_(tc|halo|stem_pair)'      35 35 ""
python scripts/ncu_traffic.py /tmp/prof/${tag}_conv.ncu-rep gpurun_out/${tag}_conv_traffic.json
run sample 'deform_sample_kernel' 4 4 "--import-source on"
run stem   'conv_stem_pair_kernel' 1 1 "--import-source on"
run post   '(nms_segment|decode_transpose|l2norm_pool)' 4 4 ""
# projection GEMM + sampler of pyramid level 0 (b32, 40x40) on their own
export LEVELS=40 CHUNKS=1024
timeout 600 ncu --set full --clock-control none -k regex:'(conv_tc_kernel|deform_sample_kernel)' -s 12 -c 2 -o /tmp/prof/${tag}_deform0 -f python scripts/bench_deform.py > /tmp/prof/ncu_deform0.log 2>&1
python scripts/ncu_table.py /tmp/prof/${tag}_deform0.ncu-rep > gpurun_out/${tag}_ncu_deform_level0.txt
python scripts/ncu_summary.py /tmp/prof/${tag}_deform0.ncu-rep >> gpurun_out/${tag}_ncu_deform_level0.txt
cp /tmp/prof/${tag}_deform0.ncu-rep gpurun_out/ 2>/dev/null
rm -f gpurun_out/${tag}_launches.csv.tmp
ls -la gpurun_out/ | head -30; du -sh gpurun_out
