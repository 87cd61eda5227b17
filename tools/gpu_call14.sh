#!/bin/bash
mkdir -p gpurun_out/r02
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_golden.py tests/test_gpu_fullsize.py -m gpu -q -k "dkr or deforconv or nofilter or families" > gpurun_out/r02/tests_call14.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02/tests_call14.log
grep -E "passed|failed|FAILED|rc=" gpurun_out/r02/tests_call14.log | tail -8
for op in fi_dkr_fwd fi_deforconv_fwd fi_nofilter_fwd; do for fl in scene up4; do python tools/run_op.py $op --flow $fl --iters 10 2>&1 | tail -1; done; done > gpurun_out/r02/fi_dkr_fwd_v1.log
cat gpurun_out/r02/fi_dkr_fwd_v1.log
