#!/bin/bash
set -x
mkdir -p gpurun_out/r02
timeout 2000 python -m pytest tests -m gpu -q --durations=15 > gpurun_out/r02/tests_call2.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02/tests_call2.log
tail -40 gpurun_out/r02/tests_call2.log
