#!/bin/bash
# last check of a round: smoke() and the whole GPU suite exactly as the driver runs them
mkdir -p gpurun_out/r02
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 1500 python -m pytest tests/ -x -q -m gpu > gpurun_out/r02/tests_final_v6.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02/tests_final_v6.log
grep -E "passed|failed|FAILED|rc=|^E  " gpurun_out/r02/tests_final_v6.log | tail -6
timeout 60 python tools/run_op.py pwc_warp --iters 20 | tail -1
