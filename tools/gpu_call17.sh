#!/bin/bash
mkdir -p gpurun_out/r02
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "tensor_core" > gpurun_out/r02/tests_call17.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02/tests_call17.log
grep -E "passed|failed|FAILED|rc=|^E  |tensor vs" gpurun_out/r02/tests_call17.log | tail -24
timeout 300 python tools/time_corr.py > gpurun_out/r02/time_corr_v1.log 2>&1
cat gpurun_out/r02/time_corr_v1.log | tail -8
