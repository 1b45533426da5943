#!/bin/bash
# Builds devit_b200/lib/libdevit_b200_trace.so: the library with the clock64 pipeline trace of
# gemm_kernel compiled in (-DDEVIT_GEMM_TRACE); select it with DEVIT_B200_LIB=<path>.
set -e
cd "$(dirname "$0")/../devit_b200/csrc"
mkdir -p build_trace
for f in common gemm attention rowops forward mlp cct edge pack; do
  [ -f $f.cu ] || continue
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo --cudart static \
       -Xcompiler -fPIC -DDEVIT_GEMM_TRACE -c $f.cu -o build_trace/$f.o &
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a --cudart static -shared \
     -o ../lib/libdevit_b200_trace.so build_trace/*.o
