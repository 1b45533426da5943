#!/bin/bash
cd "$(dirname "$0")/.."
timeout 600 python tools/time_chain.py > gpurun_out/time_chain.txt 2>&1
timeout 900 python -m pytest tests/test_headline_parity_gpu.py -x -q -s > gpurun_out/pytest_headline.txt 2>&1
timeout 300 python bench.py --no-cpu-baseline --steps 20 > gpurun_out/bench_r2_chains.json 2> gpurun_out/bench_r2_chains.err
