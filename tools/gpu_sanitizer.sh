#!/bin/bash
# compute-sanitizer on the smallest cases that reach the production kernels (one tool per gpurun call):
#   tools/gpu_sanitizer.sh memcheck|racecheck|synccheck
TOOL=${1:-memcheck}
mkdir -p gpurun_out/r02
SEL="wide or sepconv or corr_pwc"
timeout 1500 compute-sanitizer --tool $TOOL --error-exitcode 33 python -m pytest tests/test_golden.py -m gpu -q -x -k "$SEL" > gpurun_out/r02/sanitizer_${TOOL}.log 2>&1
echo "compute-sanitizer --tool $TOOL exit code: $?" >> gpurun_out/r02/sanitizer_${TOOL}.log
grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|exit code" gpurun_out/r02/sanitizer_${TOOL}.log | tail -5
