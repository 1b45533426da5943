#!/bin/bash
cd "$(dirname "$0")/.."
timeout 600 python bench.py --steps 20 > gpurun_out/bench_r2_v4_headline.json 2> gpurun_out/bench_r2_v4_headline.err
for c in c1 c2 c3 c4; do
  timeout 600 python bench.py --config $c --steps 10 --no-cpu-baseline > gpurun_out/bench_r2_v4_$c.json 2> gpurun_out/bench_r2_v4_$c.err
done
timeout 900 python -m pytest tests/test_cct_gpu.py tests/test_edge_gpu.py -x -q > gpurun_out/pytest_cct_edge.txt 2>&1
