#!/bin/bash
mkdir -p gpurun_out/r02
{
for tw in 128 144; do
  export VFIDKR_FI_STRIP_TW=$tw
  echo "== forced tile width $tw"
  timeout 40 python tools/run_op.py fi_ori_blend --flow scene --iters 20 | tail -1
  timeout 40 python tools/run_op.py fi_ori_fwd --flow scene --iters 20 --B 1 --H 2176 --W 3904 | tail -1
  timeout 40 python tools/run_op.py fi_ori_fwd --flow scene --iters 20 --B 2 --H 2176 --W 3904 | tail -1
  timeout 40 python tools/run_op.py fi_ori_fwd --flow scene --iters 50 --B 16 --H 256 --W 448 | tail -1
  timeout 40 python tools/run_op.py fi_ori_fwd --flow scene --iters 20 --B 8 --H 736 --W 1280 | tail -1
  timeout 40 python tools/run_op.py fi_ori_fwd --flow scene --iters 20 --B 8 --C 4 | tail -1
done
unset VFIDKR_FI_STRIP_TW
timeout 60 python tools/run_op.py fi_ori_fwd --flow scene --iters 5 --B 2 --C 196 | tail -1
for i in 1 2 3; do timeout 40 python tools/run_op.py corr_l4 --iters 20 | tail -1; done
} 2>&1 | tee gpurun_out/r02/strip_width_ab_v1.log
