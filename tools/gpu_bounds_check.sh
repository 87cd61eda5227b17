#!/bin/bash
mkdir -p gpurun_out/r02
P=video-frame-interpolation-based-on-deformable-kernel-region_b200
cp $P/libvfidkr_b200.so /tmp/lib_prod.so
cp $P/_build/lib_bounds.so $P/libvfidkr_b200.so
timeout 900 python tools/bounds_check_run.py > gpurun_out/r02/bounds_check_v1.log 2>&1
echo "bounds run rc=$?" | tee -a gpurun_out/r02/bounds_check_v1.log
cp /tmp/lib_prod.so $P/libvfidkr_b200.so
tail -12 gpurun_out/r02/bounds_check_v1.log
