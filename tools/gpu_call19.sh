#!/bin/bash
mkdir -p gpurun_out/r02
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_golden.py -m gpu -q -k "backward or bwd or grad or golden" > gpurun_out/r02/tests_call19.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02/tests_call19.log
grep -E "passed|failed|FAILED|rc=|^E  " gpurun_out/r02/tests_call19.log | tail -24
for fl in scene up4 gauss; do
for op in fi_ori_bwd fi_dkr_bwd; do
timeout 120 python tools/run_op.py $op --flow $fl
done; done 2>&1 | tee gpurun_out/r02/fi_bwd_v1.log
