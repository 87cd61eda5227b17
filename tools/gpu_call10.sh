#!/bin/bash
mkdir -p gpurun_out/r02
timeout 300 python tools/time_projection.py > gpurun_out/r02/time_projection_v5.log 2>&1
cat gpurun_out/r02/time_projection_v5.log
