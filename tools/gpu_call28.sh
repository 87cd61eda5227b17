#!/bin/bash
mkdir -p gpurun_out/r02
bash tools/gpu_abc.sh T128 T144 T152
