// Blocked Cholesky and triangular inverse of a matrix resident in shared memory,
// at the granularity of the DMMA fragment (8x8 sub-blocks).
//
// Replaces the unblocked dpotrf / dtrtri the reference reaches through
// jitchol / pdinv (gpy_linalg.py:77-104,219-253) for one diagonal tile.
//
// S  : n x n (n = 8 nb), row major, stride ld (ld = 8 mod 16 doubles so that the
//      LDS.128 fragment loads are bank-conflict free).  In: SPD matrix (lower
//      part read).  Out: L (lower); the strictly upper part is left untouched.
// S2 : same shape, zero-initialised by the caller.  Out: U = L^-T (upper).
// Wsm: nb x 64 doubles; Out: inverses of the 8x8 diagonal blocks of L, row major.
//
// Per 8-column step J:   (a) one warp factors the 8x8 diagonal block in registers
// (lane r owns row r; shuffles broadcast the pivot column) and inverts it;
// (b) panel  L_IJ = C_IJ W_JJ^T  - one DMMA pair per block;  (c) trailing update
// C_IK -= L_IJ L_KJ^T - one DMMA pair per block.  The inverse is then built by
// block sub-diagonals  U_{K,K+d} = -(sum_J U_KJ L_IJ^T) W_II^T  with the
// accumulator re-used directly as the A operand of the second product.
#pragma once
#include "tile_gemm.cuh"

namespace gprf {

// Lanes 0..7 hold rows 0..7 of an 8x8 SPD block in a[8].  On return lane r holds
// row r of L in a[] (upper part zeroed) and lane c holds column c of W = L^-1 in
// w[] (= row c of L^-T).  Returns 1 + index of the first non-positive pivot, or 0.
__device__ __forceinline__ int chol8_inv8(double a[8], double w[8], int lane) {
  const unsigned FULL = 0xffffffffu;
  int fail = 0;
  double dinv_own = 1.0;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const double piv = __shfl_sync(FULL, a[c], c);
    if (!(piv > 0.0) && fail == 0) fail = c + 1;
    const double dinv = rsqrt(piv);        // 75 vs 173 cycles for 1/sqrt on the critical path
    const double l = a[c] * dinv;          // lane r: L[r][c] for r > c ; lane c: sqrt(piv)
    if (lane == c) dinv_own = dinv;
#pragma unroll
    for (int c2 = c + 1; c2 < 8; ++c2) {
      const double l2 = __shfl_sync(FULL, l, c2);
      a[c2] -= l * l2;
    }
    a[c] = (lane >= c) ? l : 0.0;
  }
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    double s = (lane == r) ? 1.0 : 0.0;
#pragma unroll
    for (int m = 0; m < r; ++m) s -= __shfl_sync(FULL, a[m], r) * w[m];
    w[r] = s * __shfl_sync(FULL, dinv_own, r);
  }
  return fail;
}

// Every lane holds the whole lower triangle (A[r][c], c <= r) of an 8x8 SPD block: the factorisation
// then needs no shuffles, its critical path per pivot is rsqrt -> mul -> fma, and everything else
// (the other columns' updates, the inverse) fills the latency gaps.  Same operations in the same
// order as chol8_inv8 => the same bits.  On return A holds L (every lane), lane c < 8 holds column c
// of W = L^-1 in w[] (w[v] = W[v][c]).  Returns 1 + index of the first non-positive pivot, or 0.
__device__ __forceinline__ int chol8_full(double (&A)[8][8], double (&w)[8], int lane) {
  int fail = 0;
  double dinv[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const double piv = A[c][c];
    if (!(piv > 0.0) && fail == 0) fail = c + 1;
    dinv[c] = rsqrt(piv);
    A[c][c] = piv * dinv[c];
#pragma unroll
    for (int r = c + 1; r < 8; ++r) A[r][c] *= dinv[c];
#pragma unroll
    for (int c2 = c + 1; c2 < 8; ++c2)
#pragma unroll
      for (int r = c2; r < 8; ++r) A[r][c2] -= A[r][c] * A[c2][c];
  }
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    double s = (lane == r) ? 1.0 : 0.0;
#pragma unroll
    for (int m = 0; m < r; ++m) s -= A[r][m] * w[m];
    w[r] = s * dinv[r];
  }
  return fail;
}

// Whole-CTA routine (all threads must call).  nwarps = blockDim.x / 32.
// fail_out (shared int, pre-zeroed): 1 + global row of the first bad pivot.
__device__ __forceinline__ void smem_potrf_trtri(double* S, double* S2, double* Wsm, int ld, int nb,
                                                 int* fail_out, int row0) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int g = lane >> 2, q = lane & 3;
  int first_fail = 0;                       // warp 0: first bad pivot so far (kept in a register, stored once at the end)
  for (int J = 0; J < nb; ++J) {
    // (a) diagonal block, warp 0
    if (warp == 0) {
      double a[8], w[8];
      const int r = lane & 7;
      const double* src = S + (8 * J + r) * ld + 8 * J;
      if (lane < 8) {
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          double2 t = *reinterpret_cast<const double2*>(src + 2 * v);
          a[2 * v] = t.x;
          a[2 * v + 1] = t.y;
        }
      } else {                            // idle lanes carry an identity block through the shuffles
#pragma unroll
        for (int v = 0; v < 8; ++v) a[v] = (v == r) ? 1.0 : 0.0;
      }
      const int f = chol8_inv8(a, w, lane);
      if (f != 0 && first_fail == 0) first_fail = row0 + 8 * J + f;
      if (lane < 8) {
        double* dl = S + (8 * J + r) * ld + 8 * J;
        double* du = S2 + (8 * J + r) * ld + 8 * J;
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          *reinterpret_cast<double2*>(dl + 2 * v) = make_double2(a[2 * v], a[2 * v + 1]);
          *reinterpret_cast<double2*>(du + 2 * v) = make_double2(w[2 * v], w[2 * v + 1]);
        }
#pragma unroll
        for (int v = 0; v < 8; ++v) Wsm[J * 64 + v * 8 + r] = w[v];   // W[v][r] (lane r owns column r)
      }
    }
    __syncthreads();
    // (b) panel: L_IJ = C_IJ * W_JJ^T
    {
      const double2 b = *reinterpret_cast<const double2*>(Wsm + J * 64 + g * 8 + 2 * q);
      for (int I = J + 1 + warp; I < nb; I += nwarps) {
        double* p = S + (8 * I + g) * ld + 8 * J + 2 * q;
        const double2 a = *reinterpret_cast<const double2*>(p);
        double c0 = 0.0, c1 = 0.0;
        dmma884(c0, c1, a.x, b.x);
        dmma884(c0, c1, a.y, b.y);
        *reinterpret_cast<double2*>(p) = make_double2(c0, c1);
      }
    }
    __syncthreads();
    // (c) trailing update: C_IK -= L_IJ L_KJ^T for J < K <= I
    {
      const int m = nb - J - 1;
      const int ntask = m * (m + 1) / 2;
      for (int t = warp; t < ntask; t += nwarps) {
        int ii = (int)((sqrtf(8.0f * (float)t + 1.0f) - 1.0f) * 0.5f);
        while ((ii + 1) * (ii + 2) / 2 <= t) ++ii;
        while (ii * (ii + 1) / 2 > t) --ii;
        const int kk = t - ii * (ii + 1) / 2;
        const int I = J + 1 + ii, K = J + 1 + kk;
        const double2 a = *reinterpret_cast<const double2*>(S + (8 * I + g) * ld + 8 * J + 2 * q);
        const double2 b = *reinterpret_cast<const double2*>(S + (8 * K + g) * ld + 8 * J + 2 * q);
        double* pc = S + (8 * I + g) * ld + 8 * K + 2 * q;
        double2 c = *reinterpret_cast<const double2*>(pc);
        dmma884(c.x, c.y, -a.x, b.x);
        dmma884(c.x, c.y, -a.y, b.y);
        *reinterpret_cast<double2*>(pc) = c;
      }
    }
    __syncthreads();
  }
  // inverse by block sub-diagonals
  for (int d = 1; d < nb; ++d) {
    for (int K = warp; K + d < nb; K += nwarps) {
      const int I = K + d;
      double c0 = 0.0, c1 = 0.0;
      for (int J = K; J < I; ++J) {
        const double2 a = *reinterpret_cast<const double2*>(S2 + (8 * K + g) * ld + 8 * J + 2 * q);
        const double2 b = *reinterpret_cast<const double2*>(S + (8 * I + g) * ld + 8 * J + 2 * q);
        dmma884(c0, c1, a.x, b.x);
        dmma884(c0, c1, a.y, b.y);
      }
      const double2 wv = *reinterpret_cast<const double2*>(Wsm + I * 64 + g * 8 + 2 * q);
      double o0 = 0.0, o1 = 0.0;
      dmma884(o0, o1, c0, wv.x);
      dmma884(o0, o1, c1, wv.y);
      *reinterpret_cast<double2*>(S2 + (8 * K + g) * ld + 8 * I + 2 * q) = make_double2(-o0, -o1);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0 && first_fail != 0) {       // (volatile: the load must not be hoisted above the thread test)
    volatile int* fo = fail_out;
    if (*fo == 0) *fo = first_fail;
  }
}

}  // namespace gprf
