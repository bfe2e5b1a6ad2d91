#!/bin/bash
tag=${1:-q}
o=gpurun_out/${tag}_mt2_experiment2.txt
mkdir -p gpurun_out
(for v in "TDRN_MT2_MIN=70" "TDRN_MT2_MIN=60" "TDRN_CLUSTER=1" "TDRN_NO_MT2=1"; do echo "# $v"; env $v TDRN_TC_VERBOSE=1 python scripts/bench_conv.py conv5_1 conv4_2 2>&1 | sort | uniq -c | cut -c1-260; done) > $o 2>&1
cat $o
