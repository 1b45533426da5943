#!/bin/bash
cd "$(dirname "$0")/.."
export DEVIT_BENCH_WATCHDOG_S=150
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_r2_v8_headline_n2.json 2> gpurun_out/bench_r2_v8_headline_n2.err
echo "rc=$?" >> gpurun_out/bench_r2_v8_headline_n2.err
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 2 --steps 20 --warmup 5 --graph-multi > gpurun_out/bench_r2_v8_headline_n2_graph.json 2> gpurun_out/bench_r2_v8_headline_n2_graph.err
echo "rc=$?" >> gpurun_out/bench_r2_v8_headline_n2_graph.err
timeout 120 python -m pytest tests/test_multigpu_gpu.py -x -q -k not_current > gpurun_out/pytest_mg2.txt 2>&1
