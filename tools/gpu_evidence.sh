#!/bin/bash
# round-2 final evidence on the final tree
mkdir -p gpurun_out/r02
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/r02/tests_final_v5.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02/tests_final_v5.log
grep -E "passed|failed|FAILED|rc=|^E  " gpurun_out/r02/tests_final_v5.log | tail -6
timeout 900 python bench.py --table gpurun_out/r02/op_table_v6.jsonl > gpurun_out/r02/bench_v8.json 2> gpurun_out/r02/bench_v8.err
echo "bench rc=$?"
timeout 600 python bench.py --impl reference > gpurun_out/r02/bench_ref_v4.json 2> gpurun_out/r02/bench_ref_v4.err
echo "bench ref rc=$?"
timeout 300 python bench.py --workload 4k_stream > gpurun_out/r02/bench_4k_n1_v3.json 2> /dev/null
echo "4k rc=$?"
timeout 300 python bench.py --timed-only --steps 5 --warmup 3 > /dev/null 2>&1
echo "timed-only rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:corr_|fi_forward|projection_' -c 400 --csv --log-file gpurun_out/r02/launches_bench_v4.csv python bench.py --timed-only --steps 5 --warmup 3 > gpurun_out/r02/ncu_launches.log 2>&1
echo "ncu rc=$?"
python - <<'PY'
import json
for f in ('bench_v8','bench_ref_v4','bench_4k_n1_v3'):
    try:
        d=json.loads(open(f'gpurun_out/r02/{f}.json').read().strip().splitlines()[-1])
        print(f, {k:d.get(k) for k in ('value','ms_per_step','gpu_launches')}, (d.get('roofline') or {}).get('frac'), (d.get('e2e') or {}).get('value'), (d.get('check') or {}).get('ok'))
    except Exception as e: print(f, 'ERR', e)
PY
