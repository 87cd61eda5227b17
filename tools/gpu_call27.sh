#!/bin/bash
mkdir -p gpurun_out/r02
for fl in scene up4 gauss smooth; do timeout 60 python tools/run_op.py fi_ori_fwd --flow $fl --iters 20 2>&1 | tail -1; done | tee gpurun_out/r02/fi_fwd_v2.log
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -m gpu -q -k "fi_ or strip or blend or golden or filter" --timeout 120 > gpurun_out/r02/tests_call27.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02/tests_call27.log
grep -E "passed|failed|FAILED|rc=|^E  " gpurun_out/r02/tests_call27.log | tail -12
