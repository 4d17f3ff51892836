// Microbenchmark: does a predicated-off DMMA (mma.sync m8n8k4 f64) still occupy the FP64 tensor pipe?
//   mode 0: every DMMA executes          mode 1: odd DMMAs predicated off (@!P DMMA in SASS)
//   mode 2: dependent pairs back to back (latency probe), all executed
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
      : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int MODE>
__global__ void __launch_bounds__(128) k(double* out, int iters, int flag) {
  double c[16][2];
#pragma unroll
  for (int i = 0; i < 16; ++i) c[i][0] = c[i][1] = 0.0;
  double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-4;
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < 16; ++i) dmma(c[i][0], c[i][1], a, b);
    } else if (MODE == 1) {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        if ((i & 1) == 0 || flag) dmma(c[i][0], c[i][1], a, b);
      }
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) dmma(c[i & 1][0], c[i & 1][1], a, b);   // 2 chains only
    }
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE>
float run(double* d, int blocks, int iters, int flag) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<blocks, 128>>>(d, iters, flag); cudaDeviceSynchronize();
  cudaEventRecord(e0); k<MODE><<<blocks, 128>>>(d, iters, flag); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
  double* d; cudaMalloc(&d, 148 * 8 * 128 * 8);
  const int iters = 20000;
  for (int cps = 1; cps <= 4; cps *= 2) {
    int blocks = 148 * cps;
    float t0 = run<0>(d, blocks, iters, 1), t1 = run<1>(d, blocks, iters, 0), t1b = run<1>(d, blocks, iters, 1), t2 = run<2>(d, blocks, iters, 1);
    double fl = 2.0 * 256 * 16 * (double)iters * 4 * blocks;
    printf("CTAs/SM %d: all %.3f ms (%.1f TF/s) | half predicated off %.3f ms | same kernel flag=1 %.3f ms | 2 dependent chains %.3f ms\n",
           cps, t0, fl / t0 * 1e-9, t1, t1b, t2);
  }
  return 0;
}
