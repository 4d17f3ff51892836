// Cycle counts of the 8x8 diagonal-block routines of resident.cuh / smem_chol.cuh (one warp, shared memory).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I gprf_b200/csrc scripts/chol8_bench.cu -o scripts/chol8_bench.bin
#include <cstdio>
#include <cuda_runtime.h>
#define GPRF_RES_KERNEL_ONLY
#include "resident.cuh"
using namespace gprf;
using namespace gprf::res;

__global__ void k_bench(long long* out, int reps) {
  const int lane = threadIdx.x & 31;
  // SPD block at offset 1024, work at 2048
  for (int e = threadIdx.x; e < 64; e += 32) {
    const int r = e >> 3, c = e & 7;
    g_smem[1024 + sw_off(r, c)] = (r == c ? 10.0 : 0.0) + 1.0 / (1.0 + r + c);
  }
  if (threadIdx.x == 0) *S_FAIL = 0;
  __syncwarp();
  long long t0 = clock64();
  for (int i = 0; i < reps; ++i) {
    for (int e = lane; e < 64; e += 32) g_smem[2048 + e] = g_smem[1024 + e];
    __syncwarp();
    chol_diag_block(2048, 3072, 0, true);
    __syncwarp();
  }
  long long t1 = clock64();
  for (int i = 0; i < reps; ++i) {
    for (int e = lane; e < 64; e += 32) g_smem[2048 + e] = g_smem[1024 + e];
    __syncwarp();
    double av[8], wv[8];
    const int r = lane & 7;
    if (lane < 8) {
#pragma unroll
      for (int v = 0; v < 8; ++v) av[v] = g_smem[2048 + sw_off(r, v)];
    } else {
#pragma unroll
      for (int v = 0; v < 8; ++v) av[v] = (v == r) ? 1.0 : 0.0;
    }
    const int f = chol8_inv8(av, wv, lane);
    if (lane == 0 && f != 0) *S_FAIL = f;
    if (lane < 8) {
#pragma unroll
      for (int v = 0; v < 8; ++v) {
        g_smem[2048 + sw_off(r, v)] = av[v];
        g_smem[3072 + sw_off(v, r)] = wv[v];
      }
    }
    __syncwarp();
  }
  long long t2 = clock64();
  // rsqrt chain
  double x = g_smem[1024] + lane;
  for (int i = 0; i < reps * 8; ++i) x = rsqrt(x) + 1.5;
  long long t3 = clock64();
  double y = x;
  for (int i = 0; i < reps * 8; ++i) y = fma(y, 0.999, 0.5);
  long long t4 = clock64();
  if (threadIdx.x == 0) {
    out[0] = (t1 - t0) / reps;
    out[1] = (t2 - t1) / reps;
    out[2] = (t3 - t2) / (reps * 8);
    out[3] = (t4 - t3) / (reps * 8);
    out[4] = (long long)(x + y);
  }
}


__global__ void k_pieces(long long* out, int reps) {
  const int lane = threadIdx.x & 31;
  for (int e = threadIdx.x; e < 64; e += 32) {
    const int r = e >> 3, c = e & 7;
    g_smem[1024 + sw_off(r, c)] = (r == c ? 10.0 : 0.0) + 1.0 / (1.0 + r + c);
  }
  __syncwarp();
  long long acc[5] = {0, 0, 0, 0, 0};
  for (int i = 0; i < reps; ++i) {
    const int blk = 1024, wd = 3072, dst = 2048;
    long long t0 = clock64();
    double A[8][8], wv[8];
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int cp = 0; cp < 4; ++cp) {
        if (2 * cp > r) continue;
        const double2 t = *reinterpret_cast<const double2*>(g_smem + blk + sw_off(r, 2 * cp));
        A[r][2 * cp] = t.x;
        A[r][2 * cp + 1] = t.y;
      }
    long long t1 = clock64();
    const int f = chol8_full(A, wv, lane);
    long long t2 = clock64();
    if (lane < 8) {
#pragma unroll
      for (int v = 0; v < 8; ++v) g_smem[wd + sw_off(v, lane)] = wv[v];
    }
    long long t3 = clock64();
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int cp = 0; cp < 4; ++cp) {
        const double v0 = (2 * cp <= r) ? A[r][2 * cp] : 0.0;
        const double v1 = (2 * cp + 1 <= r) ? A[r][2 * cp + 1] : 0.0;
        *reinterpret_cast<double2*>(g_smem + dst + sw_off(r, 2 * cp)) = make_double2(v0, v1);
      }
    __syncwarp();
    long long t4 = clock64();
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int cp = 0; cp < 4; ++cp) {
        const double v0 = (2 * cp <= r) ? A[r][2 * cp] : 0.0;
        const double v1 = (2 * cp + 1 <= r) ? A[r][2 * cp + 1] : 0.0;
        if (lane == r) *reinterpret_cast<double2*>(g_smem + dst + 512 + sw_off(r, 2 * cp)) = make_double2(v0, v1);
      }
    __syncwarp();
    long long t5 = clock64();
    acc[0] += t1 - t0; acc[1] += t2 - t1; acc[2] += t3 - t2; acc[3] += t4 - t3; acc[4] += t5 - t4;
    if (f) g_smem[0] = 1.0;
  }
  if (threadIdx.x == 0) for (int k = 0; k < 5; ++k) out[k] = acc[k] / reps;
}

int main() {
  long long* d;
  cudaMalloc(&d, 64);
  cudaFuncSetAttribute(k_pieces, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  cudaFuncSetAttribute(k_bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  for (int it = 0; it < 2; ++it) k_bench<<<1, 32, 64 * 1024>>>(d, 200);
  long long h[8];
  cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
  printf("chol_diag_block %lld cycles, chol8_inv8 path %lld cycles, rsqrt+add chain %lld cycles/op, dfma chain %lld cycles/op (%s)\n", h[0], h[1], h[2],
         h[3], cudaGetErrorString(cudaGetLastError()));
  k_pieces<<<1, 32, 64 * 1024>>>(d, 200);
  cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
  printf("pieces: load %lld, factor+inverse %lld, W store %lld, L store (uniform) %lld, L store (lane == r) %lld (%s)\n", h[0], h[1], h[2], h[3], h[4],
         cudaGetErrorString(cudaGetLastError()));
  return 0;
}
