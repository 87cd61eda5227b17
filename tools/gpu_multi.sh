#!/bin/bash
# multi-GPU lines: tools/gpu_multi.sh N
N=${1:-2}
mkdir -p gpurun_out/r02
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@"; }
run --steps 10 --warmup 3 > gpurun_out/r02/bench_n$N.json 2> gpurun_out/r02/bench_n$N.err
run --workload 4k_stream --steps 3 --warmup 3 > gpurun_out/r02/bench_4k_n$N.json 2> gpurun_out/r02/bench_4k_n$N.err
run --impl reference --steps 3 --warmup 3 > gpurun_out/r02/bench_ref_n$N.json 2> gpurun_out/r02/bench_ref_n$N.err
python - <<PY
import json
for f in ("bench_n$N","bench_4k_n$N","bench_ref_n$N"):
    try:
        d=json.load(open(f"gpurun_out/r02/{f}.json")); print(f, {k:d.get(k) for k in ('value','ms_per_step','n_gpus','scaling')}, (d.get('e2e') or {}).get('value'), (d.get('e2e') or {}).get('h2d_ceiling_GBps'), (d.get('e2e') or {}).get('frac_of_h2d_ceiling'))
    except Exception as e:
        print(f, 'FAILED', e)
PY
tail -3 gpurun_out/r02/bench_4k_n$N.err
