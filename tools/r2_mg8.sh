#!/bin/bash
cd "$(dirname "$0")/.."
export DEVIT_BENCH_WATCHDOG_S=170
for c in headline c3; do
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus 8 --config $c --steps 20 --warmup 5 > gpurun_out/bench_r2_v9_${c}_n8.json 2> gpurun_out/bench_r2_v9_${c}_n8.err
  echo "rc=$?" >> gpurun_out/bench_r2_v9_${c}_n8.err
done
