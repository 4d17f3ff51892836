// FP64 tensor-core tile engine shared by every dense kernel of the path.
//
// One CTA (128 threads = 4 warps) owns one 64x64 output tile, seen as 8x8 blocks
// of 8x8 doubles (the DMMA m8n8k4 fragment).  Warp w owns block rows {w, 7-w}
// (a folded assignment: triangular operands and triangular outputs then cost
// every warp the same) and all 8 block columns: 2 x 8 accumulator fragments =
// 32 doubles per thread.  Every product on the path is of the form
//
//     C(64x64) += sum_j  A_j(64x64) * B_j(64x64)^T          ("NT")
//
// with A_j, B_j row-major tiles in HBM/L2, so a single staging pattern serves
// the Cholesky update, the triangular inverse, K^-1 = U U^T, Alpha and
// G = Alpha Alpha^T.  Operands are staged through shared memory in K-chunks
// of 32 with cp.async (16-byte, L2 only), double buffered.
//
// Block masks.  Units are padded to multiples of 64 points, diagonal tiles of
// L / U / W are triangular and diagonal output tiles are symmetric.  Executing
// those as dense 64^3 products costs 2-4x the useful flops on 100-200 point
// units, so every product carries a mask at 8x8-block granularity:
//   mlim / nlim   number of non-padding block rows / block columns of C
//   klim(j)       number of non-padding block columns of A_j, B_j
//   atri          A_j is upper triangular (block (m,k) is zero for k < m)
//   btri          B_j is upper triangular (2: block (n,k) is zero for k < n)
//   CLOW          only the lower block triangle of C is needed (n <= m)
// A skipped block leaves its accumulator fragment at zero, which is exactly the
// value of the padded / structurally-zero entry, so stores stay dense.
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
constexpr int NB8 = T / 8;       // 8x8 blocks per tile edge
constexpr int KC = 32;           // K chunk staged per pipeline stage
constexpr int SLD = KC + 8;      // smem row stride (doubles): 320 B = 64 mod 128
// Warps per CTA = warps per 64x64 tile.  4: warp w owns block rows {w, 7-w} (32 accumulator
// doubles per thread); 8: warp w owns block row w (16 per thread).  One warp can issue one
// DMMA.8x8x4 every 32 cycles while an SM sub-partition retires one every 16
// (scripts/fp64_latency.cu, profiles/r01_fp64_latency.txt), so the latency of a tile task is
// set by DMMAs per warp: 8 warps halve it and leave 4 warps per sub-partition with 2 CTAs/SM.
#ifndef GPRF_NW
#define GPRF_NW 8
#endif
constexpr int NW = GPRF_NW;
static_assert(NW == 4 || NW == 8, "4 or 8 warps per tile");
constexpr int MB = NB8 / NW;     // block rows per warp
constexpr int NTHREADS = NW * 32;
constexpr int STAGE_DOUBLES = T * SLD;               // one operand, one stage
constexpr int PIPE_DOUBLES = 4 * STAGE_DOUBLES;      // A,B x 2 stages
// Operand staging.  0 (default): per-thread 16-byte cp.async (LDGSTS), double buffered.
// 1 (-DGPRF_TMA=1): TMA bulk copies (cp.async.bulk, SASS UBLKCP) issued by one warp, one
// 256-byte row per copy into the padded stage rows, completion signalled on an mbarrier per
// stage (complete_tx) that every warp waits on, one __syncthreads per chunk instead of two.
// Correct (all GPU parity tests pass) but measured SLOWER on B200: the row-major working
// matrices allow only 256-byte copies (128 per stage), and at that granularity the copy
// engine's per-operation cost dominates - n=200k 8-rank shard 11.6 -> 19.8 ms/eval, README
// config 0.92 -> 1.28 ms.  TMA would need the tiles stored in HBM as contiguous stage images
// (one 20 KB copy per operand chunk); see DESIGN.md section 5.
#ifndef GPRF_TMA
#define GPRF_TMA 0
#endif
constexpr int PIPE_ALLOC_DOUBLES = PIPE_DOUBLES + 8; // + 3 mbarriers (stage 0, stage 1, tail tile)
constexpr int WLD = T + 8;       // stride of a fully staged 64x64 operand (72)
constexpr int SQLD = T + 1;      // stride of the scalar potf2 scratch tile

struct Acc {
  double c[MB][8][2];
};

__device__ __forceinline__ void acc_zero(Acc& a) {
#pragma unroll
  for (int m = 0; m < MB; ++m)
#pragma unroll
    for (int n = 0; n < 8; ++n) a.c[m][n][0] = a.c[m][n][1] = 0.0;
}

// Block row (0..7) of accumulator slot m for this thread's warp.
__device__ __forceinline__ int acc_brow(int m) {
  const int w = threadIdx.x >> 5;
  return (MB == 1 || m == 0) ? w : 7 - w;
}
// Accumulator element coordinates inside the 64x64 tile.
__device__ __forceinline__ int acc_row(int m) { return acc_brow(m) * 8 + ((threadIdx.x & 31) >> 2); }
__device__ __forceinline__ int acc_col(int n) { return n * 8 + 2 * (threadIdx.x & 3); }

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

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%1], %0;\n" ::"r"(count), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%1], %0;\n" ::"r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "LAB_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE;\n\t"
      "bra LAB_WAIT;\n\t"
      "DONE:\n\t"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
// global -> shared bulk copy (bytes: multiple of 16, both addresses 16-byte aligned)
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// One source tile of a product and its block mask.
struct TileRef {
  const double* p;    // element (0,0)
  long long ld;
  int klim;           // non-padding block columns (1..8)
  int atri;           // used as A: upper triangular
  int btri;           // used as B: 2 = upper triangular
};
__device__ __forceinline__ TileRef tile_ref(const double* p, long long ld, int klim = NB8, int atri = 0,
                                            int btri = 0) {
  TileRef t;
  t.p = p; t.ld = ld; t.klim = klim; t.atri = atri; t.btri = btri;
  return t;
}

// Stage the first `rows` rows of one 64 x KC chunk (row-major source, leading dimension ld).
__device__ __forceinline__ void stage_chunk(double* dst, const double* src, long long ld, int rows) {
  const int tid = threadIdx.x;
#pragma unroll
  for (int it = 0; it < (T * KC / 2) / NTHREADS; ++it) {
    int id = tid + it * NTHREADS;
    int row = id / (KC / 2);
    int cu = id % (KC / 2);
    if (row < rows) cp_async16(dst + row * SLD + cu * 2, src + (long long)row * ld + cu * 2);
  }
}

// One staged chunk: acc += A_chunk(64 x KC) * B_chunk(64 x KC)^T under the block mask.
// kk0 = block-column index of the chunk's first column inside its tile (0 or 4).
// A chunk without any masked block takes the branch-free path (all LDS.128 issued up front,
// 32 back-to-back DMMAs per 8 columns); the masked path keeps the loads unconditional and
// predicates only the DMMAs, so the scheduler can still hoist the loads.
template <bool CLOW>
__device__ __forceinline__ void mma_chunk(Acc& acc, const double* sA, const double* sB, int kk0, int mlim,
                                          int nlim, int klim, int atri, int btri) {
  const int lane = threadIdx.x & 31;
  const int g = lane >> 2, q = lane & 3;
  int br[MB];
  const double* pa[MB];
#pragma unroll
  for (int m = 0; m < MB; ++m) {
    br[m] = acc_brow(m);
    pa[m] = sA + (br[m] * 8 + g) * SLD + 2 * q;
  }
  const double* pb = sB + g * SLD + 2 * q;
  if (!CLOW && mlim == NB8 && nlim == NB8 && klim >= kk0 + KC / 8 && atri == 0 && btri == 0) {
#pragma unroll
    for (int k8 = 0; k8 < KC / 8; ++k8) {
      double2 a[MB], b[8];
#pragma unroll
      for (int m = 0; m < MB; ++m) a[m] = *reinterpret_cast<const double2*>(pa[m] + k8 * 8);
#pragma unroll
      for (int n = 0; n < 8; ++n) b[n] = *reinterpret_cast<const double2*>(pb + n * 8 * SLD + k8 * 8);
#pragma unroll
      for (int m = 0; m < MB; ++m)
#pragma unroll
        for (int n = 0; n < 8; ++n) {
          dmma884(acc.c[m][n][0], acc.c[m][n][1], a[m].x, b[n].x);
          dmma884(acc.c[m][n][0], acc.c[m][n][1], a[m].y, b[n].y);
        }
    }
    return;
  }
  // Masked path.  A predicated-off DMMA occupies the FP64 pipe exactly like an executed one
  // (measured: scripts/dmma_pred.cu, profiles/r01_dmma_predication.txt), so blocks are skipped
  // with real forward branches: per block row an ascending loop over the block columns that
  // breaks at the first masked one (every mask on the path is an upper limit on n).
#pragma unroll
  for (int k8 = 0; k8 < KC / 8; ++k8) {
    const int kk = kk0 + k8;
    if (kk >= klim) break;
    int nhi = nlim;
    if (btri == 2) nhi = min(nhi, kk + 1);
    int nh[MB];
    bool any = false;
#pragma unroll
    for (int m = 0; m < MB; ++m) {
      nh[m] = (br[m] < mlim && (!atri || kk >= br[m])) ? nhi : 0;
      if (CLOW) nh[m] = min(nh[m], br[m] + 1);
      any = any || nh[m] > 0;
    }
    if (!any) continue;
    double2 a[MB], b[8];
#pragma unroll
    for (int m = 0; m < MB; ++m) a[m] = *reinterpret_cast<const double2*>(pa[m] + k8 * 8);
#pragma unroll
    for (int n = 0; n < 8; ++n) b[n] = *reinterpret_cast<const double2*>(pb + n * 8 * SLD + k8 * 8);
#pragma unroll
    for (int m = 0; m < MB; ++m) {
#pragma unroll
      for (int n = 0; n < 8; ++n) {
        if (n >= nh[m]) break;
        dmma884(acc.c[m][n][0], acc.c[m][n][1], a[m].x, b[n].x);
        dmma884(acc.c[m][n][0], acc.c[m][n][1], a[m].y, b[n].y);
      }
    }
  }
}

__device__ __forceinline__ void stage_full_async(double* dst, const double* src, long long ld);

struct NoHook {
  __device__ __forceinline__ void operator()() const {}
};

#if GPRF_TMA
// Producer side (warp 0 only): one stage = the first arows / brows rows of an A and a B chunk.
__device__ __forceinline__ void tma_issue_chunk(double* dst, const double* pa, long long lda, const double* pb,
                                                long long ldb, int arows, int brows, uint64_t* bar) {
  const int lane = threadIdx.x;
  if (lane == 0) mbar_expect_tx(bar, (uint32_t)(arows + brows) * (KC * 8));
  __syncwarp();
  for (int r = lane; r < arows; r += 32) bulk_g2s(dst + r * SLD, pa + (long long)r * lda, KC * 8, bar);
  double* dB = dst + STAGE_DOUBLES;
  for (int r = lane; r < brows; r += 32) bulk_g2s(dB + r * SLD, pb + (long long)r * ldb, KC * 8, bar);
}
// full 64x64 row-major tile -> stride WLD
__device__ __forceinline__ void tma_issue_full(double* dst, const double* src, long long ld, uint64_t* bar) {
  const int lane = threadIdx.x;
  if (lane == 0) mbar_expect_tx(bar, (uint32_t)(T * T * 8));
  __syncwarp();
  for (int r = lane; r < T; r += 32) bulk_g2s(dst + r * WLD, src + (long long)r * ld, T * 8, bar);
}
#endif

// acc += sum_{j=0}^{nk-1} A_j * B_j^T.  tileA(j) / tileB(j) return TileRefs (klim is taken
// from A's ref and must agree with B's).  `pipe` = PIPE_DOUBLES of smem.
// Ends with a __syncthreads(): `pipe` may be reused by the caller afterwards.
//
// Two hooks hide the fixed latencies of a short contraction (units of 100-250 points have
// 0-3 tiles of K):
//   overlap()  runs right after the cp.async of the first chunk were issued (the caller
//              initialises the accumulator there: covariance values, Y^T rows); it may
//              contain __syncthreads().  Called exactly once, also when nk == 0.
//   tailW      a full 64x64 row-major tile (leading dimension tail_ld) fetched into the idle
//              stage buffer while the last chunk is multiplied; the function returns its
//              address in shared memory (stride WLD), ready to use.
#if GPRF_TMA
template <bool CLOW, class FA, class FB, class Hook = NoHook>
__device__ __forceinline__ const double* gemm_nt(Acc& acc, int nk, FA tileA, FB tileB, int mlim, int nlim,
                                                 double* pipe, Hook overlap = Hook(),
                                                 const double* tailW = nullptr, long long tail_ld = T) {
  constexpr int CPT = T / KC;        // chunks per tile
  constexpr int BPC = KC / 8;        // 8-blocks per chunk
  uint64_t* bars = reinterpret_cast<uint64_t*>(pipe + PIPE_DOUBLES);
  const bool prod = threadIdx.x < 32;          // warp 0 issues the bulk copies
  // Arm the barriers for this product: every phase of an earlier product has completed (all of
  // its copies were waited for), so the objects are simply re-initialised and parities restart
  // at 0.  The __syncthreads also orders everyone's generic accesses to `pipe` before the
  // async-proxy writes that follow.
  if (threadIdx.x == 0) {
    mbar_init(bars + 0, 1);
    mbar_init(bars + 1, 1);
    mbar_init(bars + 2, 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (prod) fence_proxy_async();
  if (nk <= 0 || mlim <= 0 || nlim <= 0) {
    if (tailW && prod) tma_issue_full(pipe, tailW, tail_ld, bars + 2);
    overlap();
    if (tailW) {
      mbar_wait(bars + 2, 0);
      __syncthreads();
    }
    return pipe;
  }
  const int arows = mlim * 8, brows = nlim * 8;
  int j = 0, h = 0;
  TileRef a = tileA(0), b = tileB(0);
  if (prod) tma_issue_chunk(pipe, a.p, a.ld, b.p, b.ld, arows, brows, bars + 0);
  overlap();
  int cur = 0;
  uint32_t par0 = 0, par1 = 0;
  const double* wsm = pipe;
  while (true) {
    int jn = j, hn = h + 1;
    if (hn == CPT || hn * BPC >= a.klim) { hn = 0; ++jn; }
    const bool more = jn < nk;
    TileRef an = a, bn = b;
    if (more) {
      if (jn != j) { an = tileA(jn); bn = tileB(jn); }
      if (prod)
        tma_issue_chunk(pipe + (cur ^ 1) * 2 * STAGE_DOUBLES, an.p + hn * KC, an.ld, bn.p + hn * KC, bn.ld, arows,
                        brows, bars + (cur ^ 1));
    } else if (tailW) {
      double* nxt = pipe + (cur ^ 1) * 2 * STAGE_DOUBLES;
      if (prod) tma_issue_full(nxt, tailW, tail_ld, bars + 2);
      wsm = nxt;
    }
    if (cur) { mbar_wait(bars + 1, par1); par1 ^= 1; }
    else { mbar_wait(bars + 0, par0); par0 ^= 1; }
    const double* cs = pipe + cur * 2 * STAGE_DOUBLES;
    mma_chunk<CLOW>(acc, cs, cs + STAGE_DOUBLES, h * BPC, mlim, nlim, a.klim, a.atri, b.btri);
    if (!more && tailW) mbar_wait(bars + 2, 0);
    __syncthreads();                 // stage `cur` is free again (next-but-one issue overwrites it)
    if (!more) break;
    j = jn; h = hn; a = an; b = bn;
    cur ^= 1;
  }
  return wsm;
}
#else
template <bool CLOW, class FA, class FB, class Hook = NoHook>
__device__ __forceinline__ const double* gemm_nt(Acc& acc, int nk, FA tileA, FB tileB, int mlim, int nlim,
                                                 double* pipe, Hook overlap = Hook(),
                                                 const double* tailW = nullptr, long long tail_ld = T) {
  constexpr int CPT = T / KC;        // chunks per tile
  constexpr int BPC = KC / 8;        // 8-blocks per chunk
  if (nk <= 0 || mlim <= 0 || nlim <= 0) {
    if (tailW) stage_full_async(pipe, tailW, tail_ld);
    overlap();
    if (tailW) {
      cp_async_wait<0>();
      __syncthreads();
    }
    return pipe;
  }
  // stage s: A at pipe + 2 s STAGE_DOUBLES, B right after it (plain arithmetic: no pointer arrays,
  // which would live in local memory once indexed with a run-time stage)
  const int arows = mlim * 8, brows = nlim * 8;
  // chunk cursor over (tile j, half h), skipping chunks beyond the tile's klim
  int j = 0, h = 0;
  TileRef a = tileA(0), b = tileB(0);
  stage_chunk(pipe, a.p, a.ld, arows);
  stage_chunk(pipe + STAGE_DOUBLES, b.p, b.ld, brows);
  cp_async_commit();
  overlap();
  int cur = 0;
  const double* wsm = pipe;
  while (true) {
    // next chunk
    int jn = j, hn = h + 1;
    if (hn == CPT || hn * BPC >= a.klim) { hn = 0; ++jn; }
    const bool more = jn < nk;
    TileRef an = a, bn = b;
    if (more) {
      if (jn != j) { an = tileA(jn); bn = tileB(jn); }
      double* nxt = pipe + (cur ^ 1) * 2 * STAGE_DOUBLES;
      stage_chunk(nxt, an.p + hn * KC, an.ld, arows);
      stage_chunk(nxt + STAGE_DOUBLES, bn.p + hn * KC, bn.ld, brows);
      cp_async_commit();
      cp_async_wait<1>();
    } else if (tailW) {
      double* nxt = pipe + (cur ^ 1) * 2 * STAGE_DOUBLES;
      stage_full_async(nxt, tailW, tail_ld);
      wsm = nxt;
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const double* cs = pipe + cur * 2 * STAGE_DOUBLES;
    mma_chunk<CLOW>(acc, cs, cs + STAGE_DOUBLES, h * BPC, mlim, nlim, a.klim, a.atri, b.btri);
    if (!more && tailW) cp_async_wait<0>();
    __syncthreads();
    if (!more) break;
    j = jn; h = hn; a = an; b = bn;
    cur ^= 1;
  }
  return wsm;
}

#endif  // GPRF_TMA

// Stage a full 64x64 row-major operand (ld) into smem with stride WLD (T * WLD doubles, which
// fit in one pipeline stage: static_assert below).  _async: issue + commit only.
__device__ __forceinline__ void stage_full_async(double* dst, const double* src, long long ld) {
  const int tid = threadIdx.x;
#pragma unroll
  for (int it = 0; it < (T * T / 2) / NTHREADS; ++it) {
    int id = tid + it * NTHREADS;
    int row = id / (T / 2);
    int cu = id % (T / 2);
    cp_async16(dst + row * WLD + cu * 2, src + (long long)row * ld + cu * 2);
  }
  cp_async_commit();
}
static_assert(T * WLD <= 2 * STAGE_DOUBLES, "a staged full tile must fit in one (A,B) stage");
__device__ __forceinline__ void stage_full(double* dst, const double* src, long long ld) {
  stage_full_async(dst, src, ld);
  cp_async_wait<0>();
  __syncthreads();
}

// out = scale * (C * W^T) with C taken from accumulator registers (used as the
// A operand, see header) and W a fully staged 64x64 row-major LOWER TRIANGULAR
// tile (stride WLD) with wlim non-padding block rows / columns.  Rows of C beyond
// mlim blocks are padding.
__device__ __forceinline__ void mul_acc_by_wt(Acc& out, const Acc& c, const double* sW, double scale, int mlim,
                                              int wlim) {
  const int lane = threadIdx.x & 31;
  const int g = lane >> 2, q = lane & 3;
  int wl[MB];
#pragma unroll
  for (int m = 0; m < MB; ++m) wl[m] = acc_brow(m) < mlim ? wlim : 0;
  acc_zero(out);
#pragma unroll
  for (int k8 = 0; k8 < 8; ++k8) {
    if (k8 >= wlim) break;
    double2 b[8];
#pragma unroll
    for (int n = k8; n < 8; ++n)                      // W[n][k] = 0 for k > n
      b[n] = *reinterpret_cast<const double2*>(sW + (n * 8 + g) * WLD + k8 * 8 + 2 * q);
#pragma unroll
    for (int m = 0; m < MB; ++m) {
#pragma unroll
      for (int n = k8; n < 8; ++n) {
        if (n >= wl[m]) break;
        dmma884(out.c[m][n][0], out.c[m][n][1], c.c[m][k8][0], b[n].x);
        dmma884(out.c[m][n][0], out.c[m][n][1], c.c[m][k8][1], b[n].y);
      }
    }
  }
  if (scale != 1.0) {
#pragma unroll
    for (int m = 0; m < MB; ++m)
#pragma unroll
      for (int n = 0; n < 8; ++n) {
        out.c[m][n][0] *= scale;
        out.c[m][n][1] *= scale;
      }
  }
}

// Store the accumulator tile row-major (16-byte stores).
__device__ __forceinline__ void acc_store(const Acc& a, double* dst, long long ld) {
#pragma unroll
  for (int m = 0; m < MB; ++m)
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      double2 v = make_double2(a.c[m][n][0], a.c[m][n][1]);
      *reinterpret_cast<double2*>(dst + (long long)acc_row(m) * ld + acc_col(n)) = v;
    }
}

// Load an accumulator tile stored by acc_store.
__device__ __forceinline__ void acc_load(Acc& a, const double* src, long long ld) {
#pragma unroll
  for (int m = 0; m < MB; ++m)
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      const double2 v = *reinterpret_cast<const double2*>(src + (long long)acc_row(m) * ld + acc_col(n));
      a.c[m][n][0] = v.x;
      a.c[m][n][1] = v.y;
    }
}

}  // namespace gprf
