#!/bin/bash
mkdir -p gpurun_out/r02
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "lowres or slice or blend or projection or mindepth" > gpurun_out/r02/tests_call16.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02/tests_call16.log
grep -E "passed|failed|FAILED|rc=|^E  " gpurun_out/r02/tests_call16.log | tail -12
bash tools/gpu_multi.sh 4
