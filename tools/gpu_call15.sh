#!/bin/bash
mkdir -p gpurun_out/r02
timeout 900 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r02/bench_ref_v1.json 2> gpurun_out/r02/bench_ref_v1.err
python -c "
import json; d=json.load(open('gpurun_out/r02/bench_ref_v1.json')); print({k:d.get(k) for k in ('impl','value','ms_per_step','n_gpus')}, d['e2e']['value'])"
timeout 900 python bench.py --workload 4k_stream --steps 3 --warmup 3 > gpurun_out/r02/bench_4k_n1.json 2> gpurun_out/r02/bench_4k_n1.err
cat gpurun_out/r02/bench_4k_n1.json; tail -3 gpurun_out/r02/bench_4k_n1.err
# ncu launch list of the bench command (after the plain run above exited 0)
python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-check > gpurun_out/r02/plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"vfidkr|corr_|fi_|projection" -c 400 --csv --log-file gpurun_out/r02/launches_bench_v1.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-check > gpurun_out/r02/ncu_bench.log 2>&1
tail -2 gpurun_out/r02/ncu_bench.log
