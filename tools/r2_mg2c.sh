#!/bin/bash
cd "$(dirname "$0")/.."
export DEVIT_BENCH_WATCHDOG_S=170
# N=1 teacher (single chain -> 2 batches in flight), then N=2 with a forced depth of 2 (exercises the per-slot communicators)
true
true
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 2 --steps 20 --warmup 5 --pipeline-depth 2 > gpurun_out/bench_r2_v13_headline_n2_pipe.json 2> gpurun_out/bench_r2_v13_headline_n2_pipe.err
echo "rc=$?" >> gpurun_out/bench_r2_v13_headline_n2_pipe.err
