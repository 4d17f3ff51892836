// FP64 peak probe for the roofline denominator: raw DMMA (m8n8k4) and DFMA issue
// rate on every SM, no memory traffic.  Prints TFLOP/s.  (MEASURED_PEAKS.json has
// no fp64 entry; bench.py also measures a cuBLAS DGEMM through torch.)
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_dmma(double* out, int iters) {
  double c[16][2];
#pragma unroll
  for (int i = 0; i < 16; ++i) c[i][0] = c[i][1] = 0.0;
  double a = threadIdx.x * 1e-3, b = threadIdx.x * 2e-3;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_dfma(double* out, int iters) {
  double c[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) c[i] = i;
  double a = 1.0000001, b = 1e-9 * threadIdx.x;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i] = fma(c[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  int sms = p.multiProcessorCount;
  double* out;
  cudaMalloc(&out, sizeof(double) * sms * 8 * 256);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int warps = 4; warps <= 32; warps *= 2) {
    int threads = warps * 32, blocks = sms * 2, iters = 20000;
    float best_m = 1e30f, best_f = 1e30f, ms;
    for (int rep = 0; rep < 4; ++rep) {
      cudaEventRecord(e0);
      k_dmma<<<blocks, threads>>>(out, iters);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      cudaEventElapsedTime(&ms, e0, e1);
      if (rep && ms < best_m) best_m = ms;
      cudaEventRecord(e0);
      k_dfma<<<blocks, threads>>>(out, iters);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      cudaEventElapsedTime(&ms, e0, e1);
      if (rep && ms < best_f) best_f = ms;
    }
    double fl_m = (double)blocks * warps * iters * 16 * 512.0;
    double fl_f = (double)blocks * threads * iters * 16 * 2.0;
    printf("{\"sms\": %d, \"warps_per_cta\": %d, \"ctas\": %d, \"dmma_tflops\": %.2f, \"dfma_tflops\": %.2f}\n",
           sms, warps, blocks, fl_m / best_m * 1e-9, fl_f / best_f * 1e-9);
  }
  return 0;
}
