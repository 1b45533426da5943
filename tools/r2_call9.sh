#!/bin/bash
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -k "proj_mlp_fused or mlp_fused" > gpurun_out/pytest_projmlp.txt 2>&1
echo "rc=$?" >> gpurun_out/pytest_projmlp.txt
for h in 4 6; do TIME_MLP_PROJ=$h timeout 120 python tools/time_mlp.py 928 1536; done > gpurun_out/time_projmlp.txt 2>&1
TIME_MLP_PROJ=4 DEVIT_B200_LIB=devit_b200/lib/libdevit_b200_trace.so timeout 120 python tools/trace_mlp.py 928 > gpurun_out/trace_projmlp.txt 2>&1
timeout 300 python bench.py --no-cpu-baseline --steps 20 > gpurun_out/bench_r2_v6.json 2> gpurun_out/bench_r2_v6.err
