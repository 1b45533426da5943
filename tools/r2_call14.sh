#!/bin/bash
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -x -q > gpurun_out/pytest_qkv16.txt 2>&1
echo "rc=$?" >> gpurun_out/pytest_qkv16.txt
echo "== EW=16" > gpurun_out/time_qkv16.txt; timeout 120 python tools/time_qkv.py >> gpurun_out/time_qkv16.txt 2>&1
echo "== EW=8" >> gpurun_out/time_qkv16.txt; DEVIT_GEMM_EW=8 timeout 120 python tools/time_qkv.py >> gpurun_out/time_qkv16.txt 2>&1
timeout 300 python bench.py --no-cpu-baseline --steps 20 > gpurun_out/bench_r2_v11.json 2> gpurun_out/bench_r2_v11.err
