#!/bin/bash
# usage: gpurun --gpus N -- "bash tools/bench_multi.sh N headline [c2 c3 c4 ...]"   (one bench.py line per config -> gpurun_out/)
cd "$(dirname "$0")/.."
export DEVIT_BENCH_WATCHDOG_S=170
N=$1; shift
for c in "$@"; do
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29581 bench.py --gpus $N --config $c --steps 20 --warmup 5 > gpurun_out/bench_multi_${c}_n$N.json 2> gpurun_out/bench_multi_${c}_n$N.err
  echo "rc=$?" >> gpurun_out/bench_multi_${c}_n$N.err
done
