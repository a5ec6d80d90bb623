#!/bin/bash
# Builds libmccbn_b200.so (sm_100a only) in-tree.  Usage: ./build.sh [extra nvcc flags]
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
$NVCC -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo \
  -Xcompiler -fPIC -Xcompiler -fvisibility=default -diag-suppress 186,550 -ccbin /usr/bin/g++ -shared \
  -o mccbn_b200/libmccbn_b200.so mccbn_b200/csrc/mccbn_b200.cu -ldl "$@"
