#!/bin/bash
mkdir -p gpurun_out/r02
timeout 900 python bench.py --steps 20 --warmup 3 --table gpurun_out/r02/op_table_v2.jsonl > gpurun_out/r02/bench_v2.json 2> gpurun_out/r02/bench_v2.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02/bench_v2.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['roofline']['frac'], d['e2e']['value'], d['check'])
PY
grep "\[op\] Corr" gpurun_out/r02/bench_v2.err
