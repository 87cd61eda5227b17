#!/bin/bash
mkdir -p gpurun_out/r02
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_golden.py tests/test_gpu_fullsize.py -m gpu -q -k "corr or config2" > gpurun_out/r02/tests_call12.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02/tests_call12.log
grep -E "passed|failed|FAILED|rc=" gpurun_out/r02/tests_call12.log | tail -8
timeout 900 python bench.py --steps 20 --warmup 3 --table gpurun_out/r02/op_table_v1.jsonl > gpurun_out/r02/bench_v1.json 2> gpurun_out/r02/bench_v1.err
cat gpurun_out/r02/bench_v1.json
grep "\[op\]" gpurun_out/r02/bench_v1.err
tail -5 gpurun_out/r02/bench_v1.err
