#!/bin/bash
# ncu --set full with source attribution of the halo-tile kernels on single layers (scripts/bench_conv.py), the capture
# DESIGN.md section 4 names as the next step for conv2_1 / conv2_2: what does the MMA issuer wait for?
# usage: gpu_profile_halo.sh <tag> [layer ...]      (default layers: conv2_1 conv2_2; one gpurun call, one GPU)
# The .ncu-rep files come back in gpurun_out/ (read here with: ncu -i <rep> --page source --csv, or --page raw --csv).
mkdir -p gpurun_out /tmp/prof
tag=${1:-r02}; shift
layers=${@:-conv2_1 conv2_2}
for l in $layers; do
    # launch 6 of the script = first timed iteration (5 warm-ups before it)
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_halo -s 5 -c 1 \
        -o /tmp/prof/${tag}_halo_$l -f python scripts/bench_conv.py $l > /tmp/prof/ncu_halo_$l.log 2>&1
    tail -n 1 /tmp/prof/ncu_halo_$l.log | cut -c1-120
    python scripts/ncu_table.py /tmp/prof/${tag}_halo_$l.ncu-rep > gpurun_out/${tag}_ncu_halo_$l.txt
    python scripts/ncu_summary.py /tmp/prof/${tag}_halo_$l.ncu-rep >> gpurun_out/${tag}_ncu_halo_$l.txt
    cp /tmp/prof/${tag}_halo_$l.ncu-rep gpurun_out/ 2>/dev/null
done
# the same layers with the epilogue compiled out of the way (TDRN_HALO_DEBUG=1), CUDA events, for the share of the epilogue
(echo "# epilogue on"; python scripts/bench_conv.py $layers; echo "# TDRN_HALO_DEBUG=1"; TDRN_HALO_DEBUG=1 python scripts/bench_conv.py $layers) > gpurun_out/${tag}_halo_epilogue_cost.txt 2>&1
cat gpurun_out/${tag}_halo_epilogue_cost.txt
