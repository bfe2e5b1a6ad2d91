#!/bin/bash
# compute-sanitizer over the kernels changed or added after profiles/r02z_sanitizer.txt: the stem pair's mid stage / builders
# (explicit st.shared / ld.shared, bias in registers), the IEEE-half MobileNet trunk operators (stem, packed-half depthwise incl. the
# row-walking form, half-operand pointwise conv, L2Norm hand-over) and one whole MobileNet forward.  usage: gpu_sanitizer_r03.sh <tag>
mkdir -p gpurun_out
tag=${1:-r03}
out=gpurun_out/${tag}_sanitizer.txt
: > $out
run() {  # $1 tool, rest: command
    tool=$1; shift
    echo "## compute-sanitizer --tool $tool $*" >> $out
    timeout 900 compute-sanitizer --tool $tool --error-exitcode 7 "$@" > /tmp/san.log 2>&1
    echo "exit code $?" >> $out
    grep -E "passed|failed|ERROR SUMMARY|RACECHECK SUMMARY|detections|Error|hazard" /tmp/san.log | cut -c1-200 | tail -n 12 >> $out
}
run memcheck python -m pytest tests/test_gpu_ops.py tests/test_gpu_tc.py -m gpu -x -q -k "half or dw or l2norm or stem"
run memcheck python scripts/sanitizer_forward.py 66 mobilenet
run racecheck python -m pytest tests/test_gpu_tc.py -m gpu -x -q -k "stem_pair and not 320-320"
run synccheck python -m pytest tests/test_gpu_tc.py -m gpu -x -q -k "stem_pair and not 320-320"
run racecheck python scripts/sanitizer_forward.py 4
cat $out
