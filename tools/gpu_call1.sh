#!/bin/bash
# GPU call 1 (round 2): wide golden fixtures from the reference kernels, the whole GPU test-suite, a bench line.
set -x
mkdir -p gpurun_out/r02
WIDE="fi_ori_wide fi_ori_wide_gauss fi_dkr_wide fi_deforconv_wide fi_nofilter_wide fi_ori_wide_c12 depthflowproj_wide flowproj_wide corr_wide_splitk corr_wide_tiled"
python oracle/make_golden.py --out gpurun_out/golden $WIDE > gpurun_out/r02/golden.log 2>&1
cp gpurun_out/golden/*.npz tests/golden/ 2>/dev/null
timeout 1500 python -m pytest tests -m gpu -q -x --durations=15 > gpurun_out/r02/tests_call1.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02/tests_call1.log
tail -30 gpurun_out/r02/tests_call1.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02/bench_call1.json 2> gpurun_out/r02/bench_call1.err
cat gpurun_out/r02/bench_call1.json
