// FP64 tensor-core tile engine shared by every dense kernel of the path.
//
// One CTA (128 threads = 4 warps) owns one 64x64 output tile.  Warp w owns
// rows 16w..16w+15 and all 64 columns: 2 x 8 DMMA (m8n8k4, f64) accumulator
// tiles = 32 doubles per thread.  Every product on the path is of the form
//
//     C(64x64) += sum_j  A_j(64x64) * B_j(64x64)^T          ("NT")
//
// with A_j, B_j row-major tiles in HBM/L2, so a single staging pattern
// serves the Cholesky update, the triangular inverse, K^-1 = U U^T, Alpha and
// G = Alpha Alpha^T.  Operands are staged through shared memory in K-chunks
// of 32 with cp.async (16-byte, L2 only), double buffered.
//
// Fragment trick: the contraction index may be visited in any order as long
// as A and B agree.  A thread loads (row g, cols 2q,2q+1) with one LDS.128 and
// feeds col 2q to DMMA #1 and col 2q+1 to DMMA #2 (k-slot q), for A and B
// alike.  With a row stride of 40 doubles every LDS.128 quarter-warp hits 32
// distinct banks.  The same (row g, cols 2q,2q+1) shape is the DMMA
// accumulator layout, so an accumulator tile can be re-used directly as the A
// operand of a following product (used for  C * W^T  in the panel kernels).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace gprf {

constexpr int T = 64;            // tile edge
constexpr int KC = 32;           // K chunk staged per pipeline stage
constexpr int SLD = KC + 8;      // smem row stride (doubles): 320 B = 64 mod 128
constexpr int NTHREADS = 128;
constexpr int STAGE_DOUBLES = T * SLD;               // one operand, one stage
constexpr int PIPE_DOUBLES = 4 * STAGE_DOUBLES;      // A,B x 2 stages
constexpr int WLD = T + 8;       // stride of a fully staged 64x64 operand (72)
constexpr int SQLD = T + 1;      // stride of the scalar potf2 scratch tile

struct Acc {
  double c[2][8][2];
};

__device__ __forceinline__ void acc_zero(Acc& a) {
#pragma unroll
  for (int m = 0; m < 2; ++m)
#pragma unroll
    for (int n = 0; n < 8; ++n) a.c[m][n][0] = a.c[m][n][1] = 0.0;
}

// D(8x8) += A(8x4) * B(4x8);  a = A[g][q], b = B[q][g], c = {C[g][2q], C[g][2q+1]}
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
      : "+d"(c0), "+d"(c1)
      : "d"(a), "d"(b));
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

// Stage one 64 x KC chunk (row-major source, leading dimension ld) into smem.
__device__ __forceinline__ void stage_chunk(double* dst, const double* src, long long ld) {
  const int tid = threadIdx.x;
#pragma unroll
  for (int it = 0; it < (T * KC / 2) / NTHREADS; ++it) {
    int id = tid + it * NTHREADS;
    int row = id / (KC / 2);
    int cu = id % (KC / 2);
    cp_async16(dst + row * SLD + cu * 2, src + (long long)row * ld + cu * 2);
  }
}

// One staged chunk: acc += A_chunk(64 x KC) * B_chunk(64 x KC)^T
__device__ __forceinline__ void mma_chunk(Acc& acc, const double* sA, const double* sB) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, q = lane & 3;
  const double* pa = sA + (warp * 16 + g) * SLD + 2 * q;
  const double* pb = sB + g * SLD + 2 * q;
#pragma unroll
  for (int k8 = 0; k8 < KC / 8; ++k8) {
    double2 a[2], b[8];
#pragma unroll
    for (int m = 0; m < 2; ++m) a[m] = *reinterpret_cast<const double2*>(pa + m * 8 * SLD + k8 * 8);
#pragma unroll
    for (int n = 0; n < 8; ++n) b[n] = *reinterpret_cast<const double2*>(pb + n * 8 * SLD + k8 * 8);
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
      for (int n = 0; n < 8; ++n) {
        dmma884(acc.c[m][n][0], acc.c[m][n][1], a[m].x, b[n].x);
        dmma884(acc.c[m][n][0], acc.c[m][n][1], a[m].y, b[n].y);
      }
  }
}

// acc += sum_{j=0}^{nk-1} A_j * B_j^T.  tileA(j) / tileB(j) return the address of
// element (0,0) of the j-th 64x64 source tile.  `pipe` = PIPE_DOUBLES of smem.
// Ends with a __syncthreads(): `pipe` may be reused by the caller afterwards.
template <class FA, class FB>
__device__ __forceinline__ void gemm_nt(Acc& acc, int nk, FA tileA, long long lda, FB tileB,
                                        long long ldb, double* pipe) {
  constexpr int CPT = T / KC;  // chunks per tile
  const int nc = nk * CPT;
  if (nc == 0) return;
  double* sA[2] = {pipe, pipe + 2 * STAGE_DOUBLES};
  double* sB[2] = {pipe + STAGE_DOUBLES, pipe + 3 * STAGE_DOUBLES};
  stage_chunk(sA[0], tileA(0), lda);
  stage_chunk(sB[0], tileB(0), ldb);
  cp_async_commit();
  for (int c = 0; c < nc; ++c) {
    const int cur = c & 1;
    if (c + 1 < nc) {
      const int j = (c + 1) / CPT, ko = ((c + 1) % CPT) * KC;
      stage_chunk(sA[cur ^ 1], tileA(j) + ko, lda);
      stage_chunk(sB[cur ^ 1], tileB(j) + ko, ldb);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    mma_chunk(acc, sA[cur], sB[cur]);
    __syncthreads();
  }
}

// Stage a full 64x64 row-major operand (ld) into smem with stride WLD.
__device__ __forceinline__ void stage_full(double* dst, const double* src, long long ld) {
  const int tid = threadIdx.x;
#pragma unroll
  for (int it = 0; it < (T * T / 2) / NTHREADS; ++it) {
    int id = tid + it * NTHREADS;
    int row = id / (T / 2);
    int cu = id % (T / 2);
    cp_async16(dst + row * WLD + cu * 2, src + (long long)row * ld + cu * 2);
  }
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();
}

// out = scale * (C * W^T) with C taken from accumulator registers (used as the
// A operand, see header) and W a fully staged 64x64 row-major tile (stride WLD).
__device__ __forceinline__ void mul_acc_by_wt(Acc& out, const Acc& c, const double* sW, double scale) {
  const int lane = threadIdx.x & 31;
  const int g = lane >> 2, q = lane & 3;
  acc_zero(out);
#pragma unroll
  for (int k8 = 0; k8 < 8; ++k8) {
    double2 b[8];
#pragma unroll
    for (int n = 0; n < 8; ++n)
      b[n] = *reinterpret_cast<const double2*>(sW + (n * 8 + g) * WLD + k8 * 8 + 2 * q);
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
      for (int n = 0; n < 8; ++n) {
        dmma884(out.c[m][n][0], out.c[m][n][1], c.c[m][k8][0], b[n].x);
        dmma884(out.c[m][n][0], out.c[m][n][1], c.c[m][k8][1], b[n].y);
      }
  }
  if (scale != 1.0) {
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
      for (int n = 0; n < 8; ++n) {
        out.c[m][n][0] *= scale;
        out.c[m][n][1] *= scale;
      }
  }
}

// Accumulator element coordinates inside the 64x64 tile.
__device__ __forceinline__ int acc_row(int m) { return (threadIdx.x >> 5) * 16 + m * 8 + ((threadIdx.x & 31) >> 2); }
__device__ __forceinline__ int acc_col(int n) { return n * 8 + 2 * (threadIdx.x & 3); }

// Store the accumulator tile row-major (16-byte stores).
__device__ __forceinline__ void acc_store(const Acc& a, double* dst, long long ld) {
#pragma unroll
  for (int m = 0; m < 2; ++m)
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      double2 v = make_double2(a.c[m][n][0], a.c[m][n][1]);
      *reinterpret_cast<double2*>(dst + (long long)acc_row(m) * ld + acc_col(n)) = v;
    }
}

}  // namespace gprf
