#!/bin/bash
# ncu --set full of the MobileNet trunk kernels of the final build: packed-half depthwise (row-walking and one-row forms), half pointwise conv
tag=${1:-q}
mkdir -p gpurun_out /tmp/prof
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dwconv -c 6 -o /tmp/prof/${tag}_dw -f python scripts/dw_only.py > /tmp/prof/ncu_dw.log 2>&1
python scripts/ncu_summary.py /tmp/prof/${tag}_dw.ncu-rep > gpurun_out/${tag}_ncu_dwconv_half.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 3 -c 1 -o /tmp/prof/${tag}_pw -f python scripts/pw_only.py 0 > /tmp/prof/ncu_pw.log 2>&1
python scripts/ncu_summary.py /tmp/prof/${tag}_pw.ncu-rep > gpurun_out/${tag}_ncu_pw512.txt 2>&1
grep -E "Kernel Name|gpu__time_duration|dram__bytes|dram_throughput|issue_active|tensor_cycles" gpurun_out/${tag}_ncu_dwconv_half.txt gpurun_out/${tag}_ncu_pw512.txt | cut -c1-200
