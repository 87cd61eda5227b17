#!/bin/bash
set -x
mkdir -p gpurun_out/r02
python tools/run_op.py dproj_fwd --flow scene --iters 2 > gpurun_out/r02/plain_dproj.log 2>&1 &&
ncu --cache-control none --clock-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active -k regex:"projection" -s 0 -c 60 --csv --log-file gpurun_out/r02/launches_dproj_chunks_warm.csv python tools/run_op.py dproj_fwd --flow scene --iters 2 > gpurun_out/r02/ncu_dproj.log 2>&1
tail -3 gpurun_out/r02/ncu_dproj.log
