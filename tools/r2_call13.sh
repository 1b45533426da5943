#!/bin/bash
# per-SM vs chip-wide memory bound of the fused kernel: same work per CTA (3 pair-tiles) on 16, 74, 148 SMs
cd "$(dirname "$0")/.."
export TIME_MLP_PROJ=4 DEVIT_B200_LIB=devit_b200/lib/libdevit_b200_trace.so
for cfg in "16 31" "74 143" "148 286"; do
  set -- $cfg
  echo "== SMs $1, images $2"
  DEVIT_SM_LIMIT=$1 TRACE_MLP_BATCH=$2 timeout 100 python tools/trace_mlp.py 928 | grep -E "us per launch|^tile|GEMM0"
done
