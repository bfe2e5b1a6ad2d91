#!/bin/bash
# compute-sanitizer over the final conv_tc.cu (early-issue epilogue, M-tile-major walk, pairing rounds rule): memcheck on the tensor-core and
# operator unit tests, racecheck / synccheck on a selection, memcheck on one MobileNet forward.   usage: gpu_sanitizer_r04.sh <tag>
mkdir -p gpurun_out
tag=${1:-r04}
out=gpurun_out/${tag}_sanitizer.txt
: > $out
run() {  # $1 tool, rest: command
    tool=$1; shift
    echo "## compute-sanitizer --tool $tool $*" >> $out
    timeout 900 compute-sanitizer --tool $tool --error-exitcode 7 "$@" > /tmp/san.log 2>&1
    echo "exit code $?" >> $out
    grep -E "passed|failed|ERROR SUMMARY|RACECHECK SUMMARY|detections|Error|hazard" /tmp/san.log | cut -c1-200 | tail -n 12 >> $out
}
run memcheck python -m pytest tests/test_gpu_tc.py tests/test_gpu_ops.py -m gpu -x -q
run memcheck python scripts/sanitizer_forward.py 8 mobilenet
run racecheck python -m pytest tests/test_gpu_tc.py tests/test_gpu_ops.py -m gpu -x -q -k "conv_tc_vs or conv1x1 or stride2 or fused_maxpool or deform_head"
run synccheck python -m pytest tests/test_gpu_tc.py tests/test_gpu_ops.py -m gpu -x -q -k "conv_tc_vs or conv1x1 or stride2 or fused_maxpool or deform_head"
cat $out
