#!/bin/bash
cd "$(dirname "$0")/.." 2>/dev/null || cd /root/repo
timeout 1400 python -m pytest tests -m gpu -q > gpurun_out/final2_pytest_gpu.txt 2>&1; echo "rc=$?" >> gpurun_out/final2_pytest_gpu.txt
DEVIT_BENCH_WATCHDOG_S=170 timeout 200 python bench.py --config c1 --steps 10 --no-cpu-baseline > gpurun_out/final2_bench_c1.json 2> gpurun_out/final2_bench_c1.err; echo "rc=$?" >> gpurun_out/final2_bench_c1.err
for c in headline c4; do
DEVIT_BENCH_WATCHDOG_S=170 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29599 bench.py --gpus 2 --config $c --steps 20 --warmup 5 --pipeline-depth 2 > gpurun_out/final2_bench_${c}_n2.json 2> gpurun_out/final2_bench_${c}_n2.err; echo "rc=$?" >> gpurun_out/final2_bench_${c}_n2.err
done
