#!/bin/bash
set -x
mkdir -p gpurun_out/r02
python tools/run_op.py dproj_fwd --flow scene --iters 2 > gpurun_out/r02/plain_dproj.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none -s 0 -c 80 --csv --log-file gpurun_out/r02/launches_dproj_chunks.csv python tools/run_op.py dproj_fwd --flow scene --iters 2 > gpurun_out/r02/ncu_dproj.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"projection_(splat|finish)_chunk|fill_mask" -s 19 -c 3 -o gpurun_out/r02/dproj_chunks_v1 python tools/run_op.py dproj_fwd --flow scene --iters 2 >> gpurun_out/r02/ncu_dproj.log 2>&1
tail -3 gpurun_out/r02/ncu_dproj.log
