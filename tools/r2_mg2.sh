#!/bin/bash
# 2-GPU checks: pytest multi-GPU tests, sharded-vs-single parity, headline + c4 bench at N=2
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_multigpu_gpu.py -x -q > gpurun_out/pytest_mg2.txt 2>&1
echo "rc=$?" >> gpurun_out/pytest_mg2.txt
CHECK_BATCH=16 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tools/check_multigpu.py > gpurun_out/check_mg2.log 2>&1
echo "rc=$?" >> gpurun_out/check_mg2.log
for c in headline c4; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --config $c --steps 20 --warmup 5 > gpurun_out/bench_r2_v8_${c}_n2.json 2> gpurun_out/bench_r2_v8_${c}_n2.err
  echo "rc=$?" >> gpurun_out/bench_r2_v8_${c}_n2.err
done
