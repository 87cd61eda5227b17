#!/bin/bash
mkdir -p gpurun_out/r02
P=video-frame-interpolation-based-on-deformable-kernel-region_b200
cp $P/libvfidkr_b200.so /tmp/lib_prod.so
cp $P/_build/libvfidkr_b200_stats.so $P/libvfidkr_b200.so
timeout 120 python tools/strip_stats.py 2>&1 | tee gpurun_out/r02/strip_stats_v1.log
cp /tmp/lib_prod.so $P/libvfidkr_b200.so
