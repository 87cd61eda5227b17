#!/bin/bash
mkdir -p gpurun_out/r02
timeout 120 python tools/run_op.py fi_dkr_fwd --flow scene --iters 1 || exit 1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fi_forward_dkr_strip -c 1 -o gpurun_out/r02/fi_dkr_r02 -f python tools/run_op.py fi_dkr_fwd --flow scene --iters 1 > gpurun_out/r02/ncu_fi_dkr.log 2>&1
echo "ncu rc=$?"
