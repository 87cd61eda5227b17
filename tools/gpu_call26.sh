#!/bin/bash
# round-2 evidence, part 2: full ncu captures (each command first run without ncu)
mkdir -p gpurun_out/r02
timeout 120 python tools/run_op.py fi_bench --iters 1 || exit 1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fi_forward_ori_strip -c 2 -o gpurun_out/r02/fi_strip_r02 -f python tools/run_op.py fi_bench --iters 1 > gpurun_out/r02/ncu_fi_strip.log 2>&1
echo "ncu fi rc=$?"
timeout 120 python tools/run_corr.py tensor 128 32 || exit 1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:corr_forward_tc -c 1 -o gpurun_out/r02/corr_tc_l5_r02 -f python tools/run_corr.py tensor 128 32 > gpurun_out/r02/ncu_corr_tc_l5.log 2>&1
echo "ncu corr rc=$?"
