#!/bin/bash
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py tests/test_headline_parity_gpu.py tests/test_cct_gpu.py -x -q > gpurun_out/pytest_split.txt 2>&1
echo "rc=$?" >> gpurun_out/pytest_split.txt
for sp in 1 0; do
  DEVIT_SPLIT_RESID=$sp timeout 300 python bench.py --no-cpu-baseline --steps 20 > gpurun_out/bench_r2_v12_split$sp.json 2> gpurun_out/bench_r2_v12_split$sp.err
done
