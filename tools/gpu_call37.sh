#!/bin/bash
mkdir -p gpurun_out/r02
{
for shape in "1 1152 1984" "2 1152 1984" "4 1152 1984" "1 736 1280" "2 736 1280" "1 2176 3904" "3 2176 3904" "1 256 448" "4 256 448"; do
  set -- $shape
  for tw in 128 144; do
    echo -n "B=$1 ${2}x$3 tw=$tw: "; VFIDKR_FI_STRIP_TW=$tw timeout 40 python tools/run_op.py fi_ori_fwd --flow scene --iters 40 --B $1 --H $2 --W $3 | tail -1
  done
done
} 2>&1 | tee gpurun_out/r02/strip_width_ab_v2.log
