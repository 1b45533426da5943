#!/bin/bash
cd "$(dirname "$0")/.."
export DEVIT_BENCH_WATCHDOG_S=170
for c in headline c2 c4; do
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29562 bench.py --gpus 4 --config $c --steps 20 --warmup 5 > gpurun_out/bench_r2_v10_${c}_n4.json 2> gpurun_out/bench_r2_v10_${c}_n4.err
  echo "rc=$?" >> gpurun_out/bench_r2_v10_${c}_n4.err
done
