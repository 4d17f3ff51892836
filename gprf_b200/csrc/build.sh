#!/bin/bash
# Build libgprf_b200.so for sm_100a (in-tree; the .so travels to the GPU box).
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
OUT=../libgprf_b200.so
$NVCC -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 \
  -Xcompiler -fPIC -shared ${GPRF_NVCC_EXTRA} \
  -o $OUT gprf_lib.cu
echo "built $(realpath $OUT)"
