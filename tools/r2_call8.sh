#!/bin/bash
cd "$(dirname "$0")/.."
export TIME_MLP_PROJ=4
echo "== 148 SMs, batch 256"; timeout 60 python tools/time_mlp.py 928
echo "== 74 SMs, batch 128 (same work per CTA)"; DEVIT_SM_LIMIT=74 TIME_MLP_BATCH=128 timeout 60 python tools/time_mlp.py 928
echo "== 74 SMs, batch 256"; DEVIT_SM_LIMIT=74 timeout 60 python tools/time_mlp.py 928
for st in 0 10000 20000 30000; do
  echo "== stagger $st"; DEVIT_MLP_STAGGER=$st timeout 60 python tools/time_mlp.py 928 1536
done
echo "== stagger mode 1, 30000"; DEVIT_MLP_STAGGER_MODE=1 DEVIT_MLP_STAGGER=30000 timeout 60 python tools/time_mlp.py 928
