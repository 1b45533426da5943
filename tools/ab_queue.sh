#!/bin/bash
# A/B experiments queued at the end of round 1 and run first thing in round 2
# (profiles/r2_ab_queue_mlp_stagger.txt).  One gpurun call, ~1 minute:
#   gpurun --timeout 300 -- 'bash tools/ab_queue.sh > gpurun_out/ab_queue.txt 2>&1'
# They decided two defaults (both kept: stagger mode 0 at 10 k clocks, token table on):
#  1. DEVIT_MLP_STAGGER_MODE=1 (delay only the clusters with a tile less) vs the current mode 0,
#     at the bs-256 shape (198 pair-tiles / 74 clusters) and at the bs-128 shape of the 8-GPU run.
#  2. DEVIT_TOK_TABLE=1 (periodic-residual patch GEMM, current default) vs 0, whole bench step.
cd "$(dirname "$0")/.."
for mode in 0 1; do
  for stag in 0 10000 25000; do
    echo "== MLP stagger mode=$mode clocks=$stag"
    DEVIT_MLP_STAGGER_MODE=$mode DEVIT_MLP_STAGGER=$stag timeout 60 python tools/time_mlp.py 928 1536
    DEVIT_MLP_STAGGER_MODE=$mode DEVIT_MLP_STAGGER=$stag TIME_MLP_BATCH=128 timeout 60 python tools/time_mlp.py 928
  done
done
for t in 1 0; do
  echo "== bench DEVIT_TOK_TABLE=$t"
  DEVIT_TOK_TABLE=$t timeout 120 python bench.py --no-cpu-baseline --steps 20 | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print(d['value'], d['ms_per_step'])"
done
