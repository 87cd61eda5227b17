#!/bin/bash
# last check of a round: smoke() and the whole GPU suite exactly as the driver runs them
mkdir -p gpurun_out/r02
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 1500 python -m pytest tests/ -x -q -m gpu > gpurun_out/r02/tests_final_v4.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02/tests_final_v4.log
grep -E "passed|failed|FAILED|rc=|^E  " gpurun_out/r02/tests_final_v4.log | tail -6
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r02/bench_v7.json 2> gpurun_out/r02/bench_v7.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02/bench_v7.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','ms_per_step','gpu_launches')}, d['roofline']['frac'], d['e2e']['value'], d['check']['ok'])
PY
