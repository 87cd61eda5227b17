#!/bin/bash
# round-2 evidence, part 1: the whole GPU suite, the bench line with the operator table, the reference arm, the launch list
mkdir -p gpurun_out/r02
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02/tests_final_v1.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02/tests_final_v1.log
grep -E "passed|failed|FAILED|rc=" gpurun_out/r02/tests_final_v1.log | tail -6
timeout 900 python bench.py --table gpurun_out/r02/op_table_v3.jsonl > gpurun_out/r02/bench_v4.json 2> gpurun_out/r02/bench_v4.err
echo "bench rc=$?"
timeout 600 python bench.py --impl reference > gpurun_out/r02/bench_ref_v2.json 2> gpurun_out/r02/bench_ref_v2.err
echo "bench ref rc=$?"
timeout 300 python bench.py --timed-only --steps 5 --warmup 3 > gpurun_out/r02/bench_timed_only.json 2> /dev/null
echo "timed-only rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02/launches_bench_v2.csv python bench.py --timed-only --steps 5 --warmup 3 > gpurun_out/r02/ncu_launches.log 2>&1
echo "ncu rc=$?"
python - <<'PY'
import json
for f in ('bench_v4','bench_ref_v2','bench_timed_only'):
    try:
        d=json.loads(open(f'gpurun_out/r02/{f}.json').read().strip().splitlines()[-1])
        print(f, {k:d.get(k) for k in ('value','ms_per_step','gpu_launches')}, (d.get('roofline') or {}).get('frac'), (d.get('e2e') or {}).get('value'), (d.get('check') or {}).get('ok'))
    except Exception as e: print(f, 'ERR', e)
PY
