#!/bin/bash
# single-GPU evidence of a round: full GPU test suite, smoke, bench (all configs), reference arm,
# ncu launch list + full capture of the dominant kernel -> gpurun_out/final_*   (gpurun -- "bash tools/final_evidence.sh")
cd "$(dirname "$0")/.."
timeout 1400 python -m pytest tests -m gpu -q > gpurun_out/final_pytest_gpu.txt 2>&1; echo "rc=$?" >> gpurun_out/final_pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.txt 2>&1; echo "rc=$?" >> gpurun_out/final_smoke.txt
timeout 400 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/final_bench_headline_n1.json 2> gpurun_out/final_bench_headline_n1.err
for c in c1 c2 c3 c4; do
  timeout 300 python bench.py --config $c --steps 10 --no-cpu-baseline > gpurun_out/final_bench_${c}_n1.json 2> gpurun_out/final_bench_${c}_n1.err
done
DEVIT_REF_BUDGET_S=20 timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_bench_reference.json 2> gpurun_out/final_bench_reference.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 2 --warmup 1 --no-graph --no-cpu-baseline --no-dense-arm > gpurun_out/final_ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mlp_fused -c 4 -f -o gpurun_out/final_prof_projmlp python tools/prof_shapes.py projmlp 2 > gpurun_out/final_ncu_projmlp.log 2>&1
