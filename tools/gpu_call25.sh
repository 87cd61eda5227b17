#!/bin/bash
mkdir -p gpurun_out/r02
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:corr_|fi_forward|projection_' -c 400 --csv --log-file gpurun_out/r02/launches_bench_v2.csv python bench.py --timed-only --steps 5 --warmup 3 > gpurun_out/r02/ncu_launches.log 2>&1
echo "ncu rc=$?"; grep -c "gpu__time_duration" gpurun_out/r02/launches_bench_v2.csv
