#!/bin/bash
set -x
mkdir -p gpurun_out/r02
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "projection or pwc_warp or config3 or lowres" > gpurun_out/r02/tests_call3_proj.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02/tests_call3_proj.log
tail -15 gpurun_out/r02/tests_call3_proj.log
timeout 600 python -m pytest tests/test_dropin_network.py tests/test_golden.py tests/test_gpu_fullsize.py -m gpu -q -s -k "network or proj" > gpurun_out/r02/tests_call3_net.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02/tests_call3_net.log
grep -E "DAIN|passed|failed|rror" gpurun_out/r02/tests_call3_net.log | tail -20
timeout 300 python tools/time_projection.py > gpurun_out/r02/time_projection_v1.log 2>&1
cat gpurun_out/r02/time_projection_v1.log
