#!/bin/bash
mkdir -p gpurun_out/r02
for fl in scene up4 gauss smooth; do python tools/run_op.py fi_ori_fwd --flow $fl --iters 20 2>&1 | tail -1; done > gpurun_out/r02/fi_fwd_v1.log
cat gpurun_out/r02/fi_fwd_v1.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02/tests_call11.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02/tests_call11.log
grep -E "passed|failed|FAILED|rc=" gpurun_out/r02/tests_call11.log | tail -12
