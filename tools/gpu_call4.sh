#!/bin/bash
set -x
mkdir -p gpurun_out/r02
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_golden.py tests/test_gpu_fullsize.py -m gpu -q -k "proj or config3 or lowres or config2" > gpurun_out/r02/tests_call4_proj.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02/tests_call4_proj.log
tail -8 gpurun_out/r02/tests_call4_proj.log
timeout 300 python tools/time_projection.py > gpurun_out/r02/time_projection_v2.log 2>&1
cat gpurun_out/r02/time_projection_v2.log
timeout 300 ./tools/microbench/_build/ffma2 > gpurun_out/r02/ffma2_microbench.log 2>&1
cat gpurun_out/r02/ffma2_microbench.log
timeout 600 python -m pytest tests/test_dropin_network.py -m gpu -q -s > gpurun_out/r02/tests_call4_net.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02/tests_call4_net.log
grep -E "256x448|passed|failed|rror" gpurun_out/r02/tests_call4_net.log | tail -8
