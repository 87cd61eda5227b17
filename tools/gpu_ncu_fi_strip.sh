#!/bin/bash
mkdir -p gpurun_out/r02
timeout 120 python tools/run_op.py fi_bench --iters 1 || exit 1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fi_forward_ori_strip -c 2 -o gpurun_out/r02/fi_strip_w144_r02 -f python tools/run_op.py fi_bench --iters 1 > gpurun_out/r02/ncu_fi_strip_w144.log 2>&1
echo "ncu fi rc=$?"
