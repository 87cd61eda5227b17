#!/bin/bash
mkdir -p gpurun_out/r02
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/r02/tests_final_v2.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02/tests_final_v2.log
grep -E "passed|failed|FAILED|rc=|^E  " gpurun_out/r02/tests_final_v2.log | tail -8
timeout 900 python bench.py --table gpurun_out/r02/op_table_v4.jsonl > gpurun_out/r02/bench_v5.json 2> gpurun_out/r02/bench_v5.err
echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02/bench_v5.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','ms_per_step','gpu_launches')}, d['roofline']['frac'], d['roofline']['frac_up4_flow'], d['roofline']['frac_iid_flow'], d['e2e']['value'], d['check']['ok'])
PY
grep -E "FI_ori_fwd|blend" gpurun_out/r02/bench_v5.err | head -12
