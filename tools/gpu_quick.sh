#!/bin/bash
# quick check after a kernel change: the tests selected by $K, then the bench line
mkdir -p gpurun_out/r02
timeout 900 python -m pytest tests -m gpu -q -k "$K" --timeout 300 2>&1 | grep -E "^E  |passed|failed" | head -8
timeout 600 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/r02/bench_quick.json 2> gpurun_out/r02/bench_quick.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02/bench_quick.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','ms_per_step','gpu_launches')}, d['roofline']['frac'], d['check']['ok'], d['check']['max_normalised_error'])
PY
