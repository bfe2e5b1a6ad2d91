#!/bin/bash
# The three compile-time experiments on the halo-tile conv kernels (conv_halo_tc.cu header), timed with scripts/bench_conv.py
# on the GPU box: each variant rebuilds only conv_halo_tc.cu with one macro, runs the layers with and without the epilogue,
# and the shipped build is restored at the end.  One gpurun call, one GPU, ~3 minutes (four nvcc runs of one file).
# usage: gpu_halo_experiments.sh <tag> [layer ...]     (default layers: conv1_2 conv2_1 conv2_2 conv3_2)
mkdir -p gpurun_out
tag=${1:-r02}; shift
layers=${@:-conv1_2 conv2_1 conv2_2 conv3_2}
out=gpurun_out/${tag}_halo_experiments.txt
: > $out
run() {   # $1 = label, $2 = nvcc macro(s) or empty
    touch tdrn_b200/csrc/conv_halo_tc.cu
    TDRN_NVCC_EXTRA="$2" python -m tdrn_b200.build > /dev/null || { echo "build failed: $1" >> $out; return; }
    echo "## $1   (epilogue on)" >> $out;  python scripts/bench_conv.py $layers >> $out 2>&1
    echo "## $1   (TDRN_HALO_DEBUG=1: epilogue skipped)" >> $out;  TDRN_HALO_DEBUG=1 python scripts/bench_conv.py $layers >> $out 2>&1
}
run "shipped build" ""
run "L2 prefetch 6 grid-strides ahead (results valid)" "-DTDRN_HALO_PREFETCH=6"
run "L2 prefetch 12 grid-strides ahead (results valid)" "-DTDRN_HALO_PREFETCH=12"
run "no A loads (TIMING ONLY)" "-DTDRN_HALO_NO_A_LOADS"
run "aligned A views (TIMING ONLY)" "-DTDRN_HALO_ALIGNED_A"
touch tdrn_b200/csrc/conv_halo_tc.cu; TDRN_NVCC_EXTRA="" python -m tdrn_b200.build > /dev/null     # back to the shipped build
cat $out
