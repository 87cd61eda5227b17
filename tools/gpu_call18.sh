#!/bin/bash
mkdir -p gpurun_out/r02
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "correlation" > gpurun_out/r02/tests_call18.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02/tests_call18.log
grep -E "passed|failed|FAILED|rc=|^E  " gpurun_out/r02/tests_call18.log | tail -24
timeout 300 python tools/time_corr.py > gpurun_out/r02/time_corr_tensor_v3.log 2>&1
tail -8 gpurun_out/r02/time_corr_tensor_v3.log
timeout 600 python bench.py > gpurun_out/r02/bench_v3.json 2> gpurun_out/r02/bench_v3.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02/bench_v3.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['roofline']['frac'], d['e2e']['value'], d.get('check'))
PY
