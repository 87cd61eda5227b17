#!/bin/bash
mkdir -p gpurun_out/r02
echo -n "first: "; VFIDKR_BWD_PLAIN=0 timeout 60 python tools/run_op.py fi_ori_bwd --flow scene || { echo "AGG kernel failed/hung rc=$?"; exit 1; }
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_golden.py -m gpu -q -x -k "backward or bwd or grad or golden" --timeout 120 > gpurun_out/r02/tests_call22.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02/tests_call22.log
grep -E "passed|failed|FAILED|rc=|^E  " gpurun_out/r02/tests_call22.log | tail -24
for fl in scene smooth up4; do
for op in fi_ori_bwd fi_dkr_bwd; do
for plain in 1 0; do
echo -n "plain=$plain "; VFIDKR_BWD_PLAIN=$plain timeout 40 python tools/run_op.py $op --flow $fl
done; done; done 2>&1 | tee gpurun_out/r02/fi_bwd_v4.log
