#!/bin/bash
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -k "mlp_fused" > gpurun_out/pytest_split.txt 2>&1
echo "rc=$?" >> gpurun_out/pytest_split.txt
for i in 1 2; do for sp in 0 1; do TIME_MLP_PROJ=4 TIME_MLP_SPLIT=$sp timeout 100 python tools/time_mlp.py 928 1536; done; done > gpurun_out/time_split.txt 2>&1
TIME_MLP_PROJ=6 TIME_MLP_SPLIT=1 timeout 100 python tools/time_mlp.py 1536 >> gpurun_out/time_split.txt 2>&1
