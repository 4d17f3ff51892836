#!/bin/bash
# Build libgprf_b200.so for sm_100a (in-tree; the .so travels to the GPU box).
# The fused per-unit kernel is compiled once per covariance family, in parallel with the
# main translation unit; objects are kept under csrc/build/ and rebuilt only when stale.
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
OUT=../libgprf_b200.so
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC ${GPRF_NVCC_EXTRA}"
mkdir -p build
HDRS="resident.cuh gprf_kernels.cuh tile_gemm.cuh smem_chol.cuh covfn.cuh partition.cuh ../../include/gprf_b200.h build.sh"
stale() {  # stale <object> <source>
  [ ! -f "$1" ] && return 0
  for f in $2 $HDRS; do [ "$f" -nt "$1" ] && return 0; done
  return 1
}
pids=()
if stale build/gprf_lib.o gprf_lib.cu; then
  $NVCC $FLAGS -c gprf_lib.cu -o build/gprf_lib.o > build/gprf_lib.log 2>&1 & pids+=($!)
fi
for D in 0 1; do for W in 0 1; do
  if stale build/gprf_fused_$D$W.o gprf_fused.cu; then
    $NVCC $FLAGS -DFUSED_DFN=$D -DFUSED_WFN=$W -c gprf_fused.cu -o build/gprf_fused_$D$W.o \
      > build/gprf_fused_$D$W.log 2>&1 & pids+=($!)
  fi
done; done
for D in 0 1; do for W in 0 1; do
  if stale build/gprf_resident_$D$W.o gprf_resident.cu; then
    $NVCC $FLAGS -DRES_DFN=$D -DRES_WFN=$W -c gprf_resident.cu -o build/gprf_resident_$D$W.o \
      > build/gprf_resident_$D$W.log 2>&1 & pids+=($!)
  fi
done; done
rc=0
for p in "${pids[@]}"; do wait $p || rc=1; done
cat build/*.log 2>/dev/null || true
[ $rc -eq 0 ] || { echo "nvcc failed"; exit 1; }
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o $OUT build/gprf_lib.o build/gprf_fused_00.o \
  build/gprf_fused_01.o build/gprf_fused_10.o build/gprf_fused_11.o build/gprf_resident_00.o \
  build/gprf_resident_01.o build/gprf_resident_10.o build/gprf_resident_11.o
echo "built $(realpath $OUT)"
