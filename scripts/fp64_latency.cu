// Microbenchmark: dependent-issue latency of the FP64 instructions the path's scalar sections
// are made of (DFMA, DMMA.8x8x4, exp(), rsqrt(), sqrt+div, shuffles), one warp per SM
// sub-partition.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/fp64_latency.bin scripts/fp64_latency.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
      : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// branch-free exp(-x) for x >= 0 (Cody-Waite reduction + degree-13 Taylor/Horner), used to
// test whether independent evaluations interleave
__device__ __forceinline__ double exp_neg(double x) {
  double t = -x;
  t = fmax(t, -708.0);
  const double n = rint(t * 1.4426950408889634);
  double r = fma(n, -6.93147180369123816490e-01, t);
  r = fma(n, -1.90821492927058770002e-10, r);
  double p = 1.0 / 6227020800.0;
  p = fma(p, r, 1.0 / 479001600.0);
  p = fma(p, r, 1.0 / 39916800.0);
  p = fma(p, r, 1.0 / 3628800.0);
  p = fma(p, r, 1.0 / 362880.0);
  p = fma(p, r, 1.0 / 40320.0);
  p = fma(p, r, 1.0 / 5040.0);
  p = fma(p, r, 1.0 / 720.0);
  p = fma(p, r, 1.0 / 120.0);
  p = fma(p, r, 1.0 / 24.0);
  p = fma(p, r, 1.0 / 6.0);
  p = fma(p, r, 0.5);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
  const int ni = (int)n;
  const double s = __hiloint2double((ni + 1023) << 20, 0);
  return p * s;
}

template <int MODE, int ILP>
__global__ void k(double* out, long long* cyc, int iters, double seed) {
  double v[ILP], w[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) { v[i] = seed + 0.001 * (threadIdx.x + i); w[i] = 0.0; }
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
      if (MODE == 0) v[i] = fma(v[i], 0.999999, 1e-9);
      if (MODE == 1) dmma884(v[i], w[i], v[i] * 1e-3, 0.5);
      if (MODE == 2) v[i] = exp(-v[i]) + 0.5;
      if (MODE == 3) v[i] = exp_neg(v[i]) + 0.5;
      if (MODE == 4) v[i] = rsqrt(v[i]) + 0.5;
      if (MODE == 5) v[i] = 1.0 / sqrt(v[i]) + 0.5;
      if (MODE == 6) v[i] = __shfl_sync(0xffffffffu, v[i], (threadIdx.x + 1) & 31);
      if (MODE == 7) v[i] = log(v[i]) + 2.0;
    }
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += v[i] + w[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int MODE, int ILP>
void run(const char* name, int threads) {
  double* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&cyc, 8);
  const int iters = 2000;
  k<MODE, ILP><<<1, threads>>>(out, cyc, iters, 1.0);
  k<MODE, ILP><<<1, threads>>>(out, cyc, iters, 1.0);
  long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-14s ILP=%2d threads=%3d : %7.1f cycles per op-group, %6.1f per op\n", name, ILP, threads,
         (double)c / iters, (double)c / iters / ILP);
  cudaFree(out); cudaFree(cyc);
}

int main() {
  run<0, 1>("dfma", 32);    run<0, 4>("dfma", 32);    run<0, 8>("dfma", 32); run<0, 8>("dfma", 128); run<0, 8>("dfma", 256);
  run<1, 1>("dmma884", 32); run<1, 4>("dmma884", 32); run<1, 16>("dmma884", 32); run<1, 16>("dmma884", 128); run<1, 16>("dmma884", 256);
  run<2, 1>("exp()", 32);   run<2, 4>("exp()", 32);   run<2, 8>("exp()", 32); run<2, 8>("exp()", 128);
  run<3, 1>("exp_neg", 32); run<3, 4>("exp_neg", 32); run<3, 8>("exp_neg", 32); run<3, 8>("exp_neg", 128); run<3, 16>("exp_neg", 128);
  run<4, 1>("rsqrt", 32);   run<4, 4>("rsqrt", 32);
  run<5, 1>("1/sqrt", 32);  run<5, 4>("1/sqrt", 32);
  run<6, 1>("shfl64", 32);  run<6, 4>("shfl64", 32);
  run<7, 1>("log", 32);     run<7, 4>("log", 32);
  return 0;
}
