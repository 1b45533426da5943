#!/bin/bash
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -k "attention" > gpurun_out/pytest_attn.txt 2>&1
echo "rc=$?" >> gpurun_out/pytest_attn.txt
timeout 120 python tools/time_attn.py > gpurun_out/time_attn.txt 2>&1
DEVIT_B200_LIB=devit_b200/lib/libdevit_b200_trace.so timeout 120 python tools/trace_attn.py 4 > gpurun_out/trace_attn.txt 2>&1
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_headline_parity_gpu.py -x -q -s > gpurun_out/pytest_model_attn.txt 2>&1
echo "rc=$?" >> gpurun_out/pytest_model_attn.txt
