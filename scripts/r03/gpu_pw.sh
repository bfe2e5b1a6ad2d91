#!/bin/bash
tag=${1:-q}
mkdir -p gpurun_out /tmp/prof
TDRN_TC_VERBOSE=1 python scripts/pw_only.py 2>&1 | sort | uniq -c | cut -c1-250 | tee gpurun_out/${tag}_pw_timing.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 3 -c 1 -o /tmp/prof/${tag}_pw512 -f python scripts/pw_only.py 0 > /tmp/prof/ncu_pw.log 2>&1
tail -n 2 /tmp/prof/ncu_pw.log | cut -c1-200
python scripts/ncu_summary.py /tmp/prof/${tag}_pw512.ncu-rep > gpurun_out/${tag}_ncu_pw512.txt 2>&1
cp /tmp/prof/${tag}_pw512.ncu-rep gpurun_out/ 2>/dev/null
head -n 40 gpurun_out/${tag}_ncu_pw512.txt | cut -c1-160
