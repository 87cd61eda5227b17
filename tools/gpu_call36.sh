#!/bin/bash
mkdir -p gpurun_out/r02
P=video-frame-interpolation-based-on-deformable-kernel-region_b200
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -k "many_channel or bigc or c196 or channels_into" --timeout 120 2>&1 | grep -E "^E  |passed|failed" | head -8
cp $P/libvfidkr_b200.so /tmp/lib_prod.so
{
for round in 1 2; do for v in NOPAIR PAIR; do
  cp $P/_build/lib_$v.so $P/libvfidkr_b200.so
  for fl in scene up4 smooth; do echo -n "variant $v round $round: "; timeout 60 python tools/run_op.py fi_ori_fwd --flow $fl --iters 5 --B 2 --C 196 | tail -1; done
done; done
} | tee gpurun_out/r02/bigc_pair_ab_v1.log
cp $P/_build/lib_bounds.so $P/libvfidkr_b200.so
timeout 900 python tools/bounds_check_run.py > gpurun_out/r02/bounds_check_v2.log 2>&1
echo "bounds run rc=$?" | tee -a gpurun_out/r02/bounds_check_v2.log
cp /tmp/lib_prod.so $P/libvfidkr_b200.so
grep -E "C>4|TOTAL" gpurun_out/r02/bounds_check_v2.log | tail -12
