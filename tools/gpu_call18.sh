#!/bin/bash
mkdir -p gpurun_out/r02
python tools/run_corr.py tensor 32 4 > gpurun_out/r02/plain_corr_tc.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:corr_forward_tc -s 2 -c 1 -o gpurun_out/r02/corr_tc_v1 python tools/run_corr.py tensor 32 4 > gpurun_out/r02/ncu_corr_tc.log 2>&1
tail -2 gpurun_out/r02/ncu_corr_tc.log
