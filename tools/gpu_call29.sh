#!/bin/bash
mkdir -p gpurun_out/r02
P=video-frame-interpolation-based-on-deformable-kernel-region_b200
cp $P/libvfidkr_b200.so /tmp/lib_prod.so
cp $P/_build/lib_D144.so $P/libvfidkr_b200.so
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -m gpu -q -x -k "fi_ or strip or blend or golden or filter or dkr" --timeout 60 2>&1 | grep -E "^E  |passed|failed" | head -8
cp /tmp/lib_prod.so $P/libvfidkr_b200.so
OP=fi_dkr_fwd bash tools/gpu_abc.sh D128 D144
for fl in scene up4 gauss smooth; do timeout 25 python tools/run_op.py fi_ori_fwd --flow $fl --iters 20 2>&1 | tail -1; done
