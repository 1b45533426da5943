#!/bin/bash
cd "$(dirname "$0")/.."
export TIME_MLP_PROJ=4
DEVIT_B200_LIB=devit_b200/lib/libdevit_b200_trace.so timeout 120 python tools/trace_mlp.py 928 > gpurun_out/trace_projmlp.txt 2>&1
unset TIME_MLP_PROJ
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mlp_fused -s 2 -c 2 -f -o gpurun_out/prof_projmlp python tools/prof_shapes.py projmlp 2 > gpurun_out/ncu_projmlp.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_persist -s 1 -c 1 -f -o gpurun_out/prof_attn python tools/prof_shapes.py attn 2 > gpurun_out/ncu_attn2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 4 -c 1 -f -o gpurun_out/prof_qkv python tools/prof_shapes.py gemm 2 > gpurun_out/ncu_qkv2.log 2>&1
