#!/bin/bash
set -x
mkdir -p gpurun_out/r02
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_golden.py tests/test_gpu_fullsize.py -m gpu -q -k "proj or config3 or lowres or config2 or mindepth" > gpurun_out/r02/tests_call5_proj.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02/tests_call5_proj.log
tail -8 gpurun_out/r02/tests_call5_proj.log
timeout 300 python tools/time_projection.py > gpurun_out/r02/time_projection_v3.log 2>&1
cat gpurun_out/r02/time_projection_v3.log
