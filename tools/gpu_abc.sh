#!/bin/bash
# A/B/C timing of library variants in ONE call (box-to-box noise is ~1-2 %): _build/lib_<X>.so
mkdir -p gpurun_out/r02
P=video-frame-interpolation-based-on-deformable-kernel-region_b200
cp $P/libvfidkr_b200.so /tmp/lib_prod.so
OP=${OP:-fi_ori_fwd}
for round in 1 2; do
for v in "$@"; do
  cp $P/_build/lib_$v.so $P/libvfidkr_b200.so
  for fl in scene up4; do
    echo -n "variant $v round $round: "; timeout 25 python tools/run_op.py $OP --flow $fl --iters 30 2>&1 | tail -1
  done
done; done | tee gpurun_out/r02/abc.log
cp /tmp/lib_prod.so $P/libvfidkr_b200.so
