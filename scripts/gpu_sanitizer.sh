#!/bin/bash
# compute-sanitizer over the hand-written kernels: memcheck on the tensor-core and post-processing unit
# tests and on one multi-stream b32 forward; racecheck and synccheck on the post-processing tests, a selection of the
# tensor-core tests and a b4 forward.  One GPU, ~10 minutes.  usage: gpu_sanitizer.sh <tag>
mkdir -p gpurun_out
tag=${1:-r02}
out=gpurun_out/${tag}_sanitizer.txt
: > $out
run() {  # $1 tool, rest: command
    tool=$1; shift
    echo "## compute-sanitizer --tool $tool $*" >> $out
    timeout 900 compute-sanitizer --tool $tool --error-exitcode 7 "$@" > /tmp/san.log 2>&1
    echo "exit code $?" >> $out
    grep -E "passed|failed|ERROR SUMMARY|RACECHECK SUMMARY|detections|Error|hazard" /tmp/san.log | cut -c1-200 | tail -n 12 >> $out
}
run memcheck python -m pytest tests/test_gpu_tc.py tests/test_gpu_postprocess.py -m gpu -x -q
run memcheck python scripts/sanitizer_forward.py 32
run racecheck python -m pytest tests/test_gpu_postprocess.py -m gpu -x -q
run racecheck python -m pytest tests/test_gpu_tc.py -m gpu -x -q -k "fused_maxpool or stride2 or split3 or deform_head or deform_sample_group or dwpw or stem_split"
run racecheck python scripts/sanitizer_forward.py 4
run synccheck python -m pytest tests/test_gpu_postprocess.py -m gpu -x -q
run synccheck python -m pytest tests/test_gpu_tc.py -m gpu -x -q -k "fused_maxpool or stride2 or split3 or deform_head or deform_sample_group or dwpw or stem_split"
run synccheck python scripts/sanitizer_forward.py 4
cat $out
