#!/bin/bash
mkdir -p gpurun_out/r02
for agg in 0 1; do
  VFIDKR_BWD_AGG=$agg timeout 120 python tools/run_op.py fi_ori_bwd --flow scene
done 2>&1 | tee gpurun_out/r02/fi_bwd_v2.log
for agg in 0 1; do
  VFIDKR_BWD_AGG=$agg timeout 600 ncu --set full --clock-control none --import-source on -k regex:fi_backward -c 1 -o gpurun_out/r02/fi_bwd_agg$agg -f python tools/run_op.py fi_ori_bwd --flow scene --iters 1 > gpurun_out/r02/ncu_fi_bwd_agg$agg.log 2>&1
  echo "ncu agg=$agg rc=$?"
done
