#!/bin/bash
cd "$(dirname "$0")/.."
export DEVIT_BENCH_WATCHDOG_S=170
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29591 bench.py --gpus 8 --steps 20 --warmup 5 --pipeline-depth 3 --no-dense-arm > gpurun_out/bench_r2_v14_headline_n8_d3.json 2> gpurun_out/bench_r2_v14_headline_n8_d3.err
echo "rc=$?" >> gpurun_out/bench_r2_v14_headline_n8_d3.err
