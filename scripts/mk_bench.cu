// Throughput of the resident kernel's block-product loop (resident.cuh: mk / mk_loop) on one SM:
// 16 warps, operands in shared memory, cycles per DMMA per sub-partition (16 = the pipe's peak).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I gprf_b200/csrc scripts/mk_bench.cu -o scripts/mk_bench.bin
#include <cstdio>
#include <cuda_runtime.h>
#define GPRF_RES_KERNEL_ONLY
#include "resident.cuh"
using namespace gprf;
using namespace gprf::res;

// variant 0: mk as in resident.cuh; 1: the two DMMAs of a pair split across the j loop (a.x for all j, then a.y)
template <bool AT, bool BT, int NJ, int VAR>
__device__ __forceinline__ void loop(double2 (&acc)[4], int pa, int sa, const int (&pb)[4], int sbk, int k1, const Lane& L) {
  if (VAR == 0) {
    mk_loop<AT, BT, NJ>(acc, pa, sa, pb, sbk, 0, k1, L);
  } else {
    unsigned a0 = (AT ? L.bt0 : L.bn) + 8u * pa, a1 = L.bt1 + 8u * pa;
    unsigned b0[NJ], b1[NJ];
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      b0[j] = (BT ? L.bt0 : L.bn) + 8u * pb[j];
      b1[j] = L.bt1 + 8u * pb[j];
    }
    const unsigned sab = 8u * sa, sbb = 8u * sbk;
#pragma unroll 1
    for (int k = 0; k < k1; ++k) {
      double2 av;
      if (AT) { av = make_double2(lds64(a0), lds64(a1)); a1 += sab; } else av = lds128(a0);
      a0 += sab;
      double2 bv[NJ];
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        if (BT) { bv[j] = make_double2(lds64(b0[j]), lds64(b1[j])); b1[j] += sbb; } else bv[j] = lds128(b0[j]);
        b0[j] += sbb;
      }
#pragma unroll
      for (int j = 0; j < NJ; ++j) dmma884(acc[j].x, acc[j].y, av.x, bv[j].x);
#pragma unroll
      for (int j = 0; j < NJ; ++j) dmma884(acc[j].x, acc[j].y, av.y, bv[j].y);
    }
  }
}

template <bool AT, bool BT, int NJ, int VAR>
__global__ void __launch_bounds__(512, 1) k_mk(long long* out, double* sink, int reps, int k1, int nwarps) {
  const Lane L = make_lane();
  for (int e = threadIdx.x; e < 400 * 64; e += blockDim.x) g_smem[e] = 1e-3 * (e % 97);
  __syncthreads();
  double2 acc[4];
  zero4(acc);
  const int pa = (L.w * 13) * 64 % (200 * 64);
  int pb[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) pb[j] = 200 * 64 + ((L.w + j * 3) % 13) * 64;
  __syncthreads();
  const long long t0 = clock64();
  if (L.w < nwarps)
    for (int r = 0; r < reps; ++r) loop<AT, BT, NJ, VAR>(acc, pa, 64, pb, 13 * 64, k1, L);
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) out[0] = t1 - t0;
  sink[threadIdx.x] = acc[0].x + acc[1].y + acc[2].x + acc[3].y;
}

template <bool AT, bool BT, int NJ, int VAR>
void run(const char* name, int nwarps) {
  long long* d; double* s;
  cudaMalloc(&d, 64); cudaMalloc(&s, 4096);
  auto kern = k_mk<AT, BT, NJ, VAR>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  const int reps = 200, k1 = 13;
  kern<<<1, 512, 220 * 1024>>>(d, s, reps, k1, nwarps);
  kern<<<1, 512, 220 * 1024>>>(d, s, reps, k1, nwarps);
  long long c; cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
  const double dmma_per_smsp = (double)reps * k1 * NJ * 2 * nwarps / 4.0;
  printf("%-34s warps %2d : %7.1f cycles per DMMA per sub-partition (%s)\n", name, nwarps, c / dmma_per_smsp, cudaGetErrorString(cudaGetLastError()));
  cudaFree(d); cudaFree(s);
}

int main() {
  for (int nw : {4, 8, 16}) {
    run<false, false, 4, 0>("A n, B n, NJ 4, pairs back to back", nw);
    run<false, false, 4, 1>("A n, B n, NJ 4, halves interleaved", nw);
    run<false, true, 4, 0>("A n, B t, NJ 4, pairs back to back", nw);
    run<false, true, 4, 1>("A n, B t, NJ 4, halves interleaved", nw);
    run<true, true, 4, 0>("A t, B t, NJ 4, pairs back to back", nw);
    run<true, true, 4, 1>("A t, B t, NJ 4, halves interleaved", nw);
    run<false, true, 2, 0>("A n, B t, NJ 2, pairs back to back", nw);
    run<false, true, 2, 1>("A n, B t, NJ 2, halves interleaved", nw);
  }
  return 0;
}
