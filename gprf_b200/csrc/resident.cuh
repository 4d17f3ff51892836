// Resident unit kernel: one CTA evaluates one unit (a block, or an edge's pair of blocks) of the
// GPRF objective (gprf.py:206-330, 496-591) entirely out of shared memory.
//
// The tile pipeline of gprf_kernels.cuh keeps every unit's working matrix in HBM/L2 and runs one
// launch per dependency level; for the README configuration (442 units of 100-250 points) that
// is a chain of ~25 latency-bound launches.  Here a unit never leaves the SM:
//
//   block unit b      K_bb -> L_b -> W_b = L_b^-1 (in place), Z_b = W_b Y_b, alpha_b = W_b^T Z_b,
//                     K_bb^-1 = W_b^T W_b, G = alpha alpha^T - dy K^-1 contracted with dK in
//                     registers; W_b, Z_b, alpha_b, K_bb^-1, logdet_b and |Z_b|^2 are exported.
//   pair unit (i, j)  the edge factorisation REUSES block i's factor and factors only the Schur
//                     complement of block j (rows of i come first, gprf.py:310-330):
//                       L_ji = K_ji W_i^T                 S = K_jj + nv I - L_ji L_ji^T
//                       L_S = chol(S), W_S = L_S^-1       Z_j = W_S (Y_j - L_ji Z_i)
//                       V = -W_S L_ji W_i                 (L^-1 of the pair = [[W_i, 0], [V, W_S]])
//                       alpha = [alpha_i + V^T Z_j ; W_S^T Z_j]
//                       K^-1 = [[K_ii^-1 + V^T V, .], [W_S^T V, W_S^T W_S]]
//                     (checked against the oracle on the CPU in tests/test_schur_model.py).
//
// Storage: matrices are kept as packed 8x8 blocks (the DMMA m8n8k4 fragment), 512 contiguous
// bytes each, with an in-block swizzle that makes BOTH the row-major fragment load (LDS.128) and
// the transposed fragment load (2 x LDS.64) bank-conflict free, so that every product on the
// path reads its operands where they already are:
//   R1  bb x ab blocks   L_ji -> T = L_ji W_i -> V         (pair units only)
//   R2  lower triangle of bb x bb blocks   S -> L_S -> W_S
//   XS  coordinate records of the unit's points, RING  3 x 8 KB staging ring
// Operands that live in HBM/L2 (the parent block's exports, this CTA's own Z / alpha scratch)
// are streamed through RING by TMA bulk copies (cp.async.bulk + mbarrier complete_tx).
// Padding rows/columns (sizes are rounded up to 8 separately for block i and block j) form an
// identity block, as in the tile pipeline.
//
// Limits: every block has at most 8*BMAXB = 160 points and dy <= 64.  A pair whose R1 does not fit in
// shared memory next to R2 (res_class 1: roughly two blocks of more than 120 points) keeps R1 in this
// CTA's L2-resident scratch instead - same code, generic loads; those few units are queued first.
// Structures outside the limits, and evaluations in which a pivot fails (the jitter rule of
// gpy_linalg.py:77-97), are reported in the status word and re-run through the tile pipeline.
#pragma once
#include "covfn.cuh"
#include "tile_gemm.cuh"
#include "smem_chol.cuh"

namespace gprf {
namespace res {

constexpr int RMAXB = 16;                 // 8x8 blocks per side of one block of points, register-array size
constexpr int BMAXB = 20;                 // ... of the "big unit" instantiation (up to 160 points per block)
constexpr int EMAXB = BMAXB;              // bound used by the exported / scratch layouts
constexpr int RNYB = 8;                   // y blocks per alpha row in the exported layouts (dy <= 64)
constexpr int RNW = 16;                   // warps per CTA
constexpr int RNT = RNW * 32;
constexpr int RBLK = 64;                  // doubles per 8x8 block
constexpr int R_WB_BLOCKS = 60;           // fixed work buffer (30 KB): diagonal inverses, right-hand sides, task table
constexpr int R_MISC_DOUBLES = 448;      // 7 blocks
constexpr int R_SMEM_BYTES = 232448;      // 227 KB, the sm_100 per-CTA maximum
// doubles left for XS (coordinate records), R1, R2 and the staging area F behind them
constexpr int R_CAP_DOUBLES = R_SMEM_BYTES / 8 - R_WB_BLOCKS * RBLK - R_MISC_DOUBLES;
constexpr int R_MIN_STAGE_BLOCKS = 32;    // F must at least hold one exported row (EMAXB blocks) / GCOLS alpha rows

constexpr int RTRI = EMAXB * (EMAXB + 1) / 2;
// per-block export (doubles); nb = the block's own number of 8-blocks
constexpr long long EXP_W = 0;                                       // lower packed: rtri(k) + c
constexpr long long EXP_KINV = (long long)RTRI * RBLK;               // lower packed
constexpr long long EXP_KSAVE = 2LL * RTRI * RBLK;                   // lower packed: covariance values
constexpr long long EXP_ZY = 3LL * RTRI * RBLK;                      // yb * nb + k   (compact)
constexpr long long EXP_AROW = EXP_ZY + (long long)RNYB * EMAXB * RBLK;     // k * nyb + yb  (compact)
constexpr long long EXP_SCAL = EXP_AROW + (long long)EMAXB * RNYB * RBLK;   // logdet, |Z|^2
constexpr long long EXP_STRIDE = EXP_SCAL + 16;
// per-CTA scratch (doubles)
constexpr long long SCR_ZY = 0;                                              // yb * bb + k
constexpr long long SCR_AROW = (long long)RNYB * EMAXB * RBLK;              // 2*EMAXB rows x RNYB
constexpr long long SCR_COLP = SCR_AROW + 2LL * EMAXB * RNYB * RBLK;
constexpr int COLP = 32;                                                     // per G block: 8 columns x 4 raw sums
constexpr long long SCR_TASKP = SCR_COLP + (long long)(2 * EMAXB) * (2 * EMAXB + 1) / 2 * COLP;
constexpr int TASKP = 72;                                                    // per task: 8 rows x 8 raw sums, 2 scalars
constexpr int GCOLS = 4;                                                     // G blocks per task
constexpr int MAXG = 24;                                                     // tasks per block row, at most
constexpr long long SCR_KJI = SCR_TASKP + (long long)(2 * EMAXB) * MAXG * TASKP;   // saved K_ji: row * ab + c
constexpr long long SCR_KJJ = SCR_KJI + (long long)EMAXB * EMAXB * RBLK;           // saved K_jj: lower packed
constexpr long long SCR_R1 = SCR_KJJ + (long long)RTRI * RBLK;                      // R1 of units too big for smem
constexpr long long SCR_STRIDE = SCR_R1 + (long long)EMAXB * EMAXB * RBLK;
constexpr int GX_STRIDE = 2 * EMAXB * 8 * 3;           // per-unit gradX rows (padded local order)

enum { ST_OVERFLOW = 1, ST_NOTPD = 2, ST_TIMEOUT = 4 };
// Watchdog of the kernel's spin waits (a parent block's flag, a TMA's mbarrier): after ~2 s (a launch takes
// 0.5 ms) the wait gives up and sets ST_TIMEOUT - the evaluation's results are then discarded and it is
// redone by the tile pipeline, like a structure that does not fit.  It turns any scheduling or
// synchronisation fault (e.g. two resident launches on one GPU whose CTAs wait for each other's SMs)
// from a hung process into a visible fallback (gprf_resident_stats).
// The limit travels in ResParams::spin_limit (cycles; GPRF_RES_WATCHDOG_S seconds in the environment, for runs
// under compute-sanitizer, where a launch takes minutes).
constexpr long long SPIN_LIMIT_CYCLES = 4000000000LL;

__host__ __device__ __forceinline__ int rtri(int i) { return i * (i + 1) / 2; }
// shared memory, in whole 8x8 blocks:  MISC | XS | R1 | R2 | F | WB
// XS: one block per 8 points (48 doubles of coordinate records; the gradient phase rebuilds it as 8 x 8 feature blocks)
__host__ __device__ __forceinline__ int res_xs_blocks(int ab, int bb) { return ab + bb; }
// free staging blocks F between R2 and WB (negative: does not fit)
__host__ __device__ __forceinline__ int res_free_blocks(int ab, int bb, bool r1_global) {
  return R_CAP_DOUBLES / RBLK - res_xs_blocks(ab, bb) - (r1_global ? 0 : bb * ab) - rtri(bb);
}
// 0: everything in shared memory; 1: R1 (the bb x ab coupling matrix) in this CTA's L2-resident
// scratch, the rest in shared memory; 2: does not fit (tile pipeline).
__host__ __device__ __forceinline__ int res_class(int ab, int bb) {
  if (ab > BMAXB || bb > BMAXB) return 2;
  const int nf = res_free_blocks(ab, bb, false);
  if (nf >= R_MIN_STAGE_BLOCKS && rtri(ab) <= nf + rtri(bb)) return 0;
  return 1;
}

struct ResParams {
  const double* X;              // n x dx
  const double* Y;              // n x dy
  const long long* perm;
  const long long* block_ptr;   // B + 1
  const int* edges;             // 2 E
  const int* order;             // units of this launch (block id, or B + edge id)
  const int* n_order;           // their number (device: written by k_res_plan)
  const int* list_ptr;          // gridDim.x + 1: CTA w evaluates order[list_ptr[w] .. list_ptr[w + 1])
  int B, dx, dy, nyb;
  int want_grad;
  CovParams cp;
  double* exports;              // B x EXP_STRIDE
  double* scratch;              // gridDim.x x SCR_STRIDE
  double* ll_u;                 // per unit
  double* gth_u;                // per unit x MAX_NCOV
  double* gx_u;                 // per unit x GX_STRIDE
  int* info;                    // per unit: 1 + first failing local row
  int* status;
  long long spin_limit;         // watchdog of the spin waits, in SM cycles
  int defer_ok;                 // a pair whose parent is still running starts with its own K_ji (GPRF_RES_DEFER)
  int* ready;                   // per block: == epoch once the block unit's factor exports (W, Z, alpha, K, scalars) are complete
  int* ready2;                  // per block: == epoch once K^-1 (written by the block's gradient phase) is complete too
  int* ready0;                  // per block: == epoch once W = L^-1 and the saved covariance values are exported (all a pair
                                // needs for L_ji, S and its factorisation); nullptr: pairs start on `ready`
  int epoch;
  int dbg_unit, dbg_phase;      // debug dump of R1 / R2 after a phase (-1: off)
  double* dbg_out;              // 2 x (160 x 160) doubles
  unsigned long long* trace;    // debug timeline (RTRACE_SLOTS (tag, ns) pairs per CTA) or nullptr
};

// ---- swizzled 8x8 block ---------------------------------------------------------------------
// element (r, c) lives at  8 * (r ^ ((r >> 1) & 1)) + (c ^ (r & 4)).
__host__ __device__ __forceinline__ int sw_off(int r, int c) { return ((r ^ ((r >> 1) & 1)) << 3) + (c ^ (r & 4)); }

struct Lane {
  int w, lane, g, q;
  int on, ot0, ot1;             // offsets (doubles) of the row-major / transposed fragment elements
  unsigned bn, bt0, bt1;        // the same as byte addresses in the shared window (base of g_smem included)
};
extern __shared__ __align__(128) double g_smem[];
__device__ __forceinline__ Lane make_lane() {
  Lane L;
  L.w = threadIdx.x >> 5;
  L.lane = threadIdx.x & 31;
  L.g = L.lane >> 2;
  L.q = L.lane & 3;
  L.on = sw_off(L.g, 2 * L.q);
  L.ot0 = sw_off(2 * L.q, L.g);
  L.ot1 = sw_off(2 * L.q + 1, L.g);
  const unsigned base = smem_u32(g_smem);
  L.bn = base + 8u * L.on;
  L.bt0 = base + 8u * L.ot0;
  L.bt1 = base + 8u * L.ot1;
  // opaque to the optimiser: under register pressure it would otherwise recompute these from
  // %tid at every use (measured: 13 % of all executed instructions)
  asm volatile("" : "+r"(L.on), "+r"(L.ot0), "+r"(L.ot1), "+r"(L.g), "+r"(L.q));
  asm volatile("" : "+r"(L.bn), "+r"(L.bt0), "+r"(L.bt1));
  return L;
}
// Shared-memory operands are addressed by their offset (in doubles) into the CTA's dynamic shared
// memory.  The fragment loads are explicit ld.shared on 32-bit window addresses, lane base + 8 * offset:
// ONE integer instruction per load.  (Indexing g_smem cost an add and a multiply-add per load - the
// window base is a run-time value on sm_100 - and the product loops ran at 11 instructions per DMMA,
// issue-bound; generic pointers, as handed around in a struct, cost a 64-bit add chain and the slower
// generic LD.)  Operands in HBM / L2 use the pointer overloads.
__device__ __forceinline__ double2 lds128(unsigned a) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ double lds64(unsigned a) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void sts128(unsigned a, double2 v) {
  asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(a), "d"(v.x), "d"(v.y) : "memory");
}
__device__ __forceinline__ double2 ldn(int off, const Lane& L) { return lds128(L.bn + 8u * (unsigned)off); }
__device__ __forceinline__ double2 ldt(int off, const Lane& L) {
  return make_double2(lds64(L.bt0 + 8u * (unsigned)off), lds64(L.bt1 + 8u * (unsigned)off));
}
__device__ __forceinline__ void stn(int off, const Lane& L, double2 v) { sts128(L.bn + 8u * (unsigned)off, v); }
// row-major fragment: (M[g][2q], M[g][2q+1])  - A operand of C = A B^T, B operand given as [n][k],
// and the accumulator layout
__device__ __forceinline__ double2 ldn(const double* blk, const Lane& L) {
  return *reinterpret_cast<const double2*>(blk + L.on);
}
// transposed fragment: (M[2q][g], M[2q+1][g]) - operand given as [k][m] / [k][n]
__device__ __forceinline__ double2 ldt(const double* blk, const Lane& L) {
  return make_double2(blk[L.ot0], blk[L.ot1]);
}
__device__ __forceinline__ void stn(double* blk, const Lane& L, double2 v) {
  *reinterpret_cast<double2*>(blk + L.on) = v;
}
__device__ __forceinline__ void mma2(double2& c, double2 a, double2 b) {
  dmma884(c.x, c.y, a.x, b.x);
  dmma884(c.x, c.y, a.y, b.y);
}
__device__ __forceinline__ double2 neg2(double2 v) { return make_double2(-v.x, -v.y); }

// f(integral_constant<K>) for K = n-1, n-2, ..., 0: one indexed jump into straight-line code, so that
// loops whose register arrays need compile-time indices run without a branch per element and the
// compiler can hoist the fragment loads of a whole pass in front of its DMMAs.
template <int K>
struct IC {
  static constexpr int value = K;
};
#define RES_CASE(K) \
  case K + 1:       \
    if constexpr (K < MB) f(IC<K>{});
template <int MB, class F>
__device__ __forceinline__ void for_desc(int n, F&& f) {
  switch (n) {
    RES_CASE(19) RES_CASE(18) RES_CASE(17) RES_CASE(16) RES_CASE(15) RES_CASE(14) RES_CASE(13) RES_CASE(12)
    RES_CASE(11) RES_CASE(10) RES_CASE(9) RES_CASE(8) RES_CASE(7) RES_CASE(6) RES_CASE(5) RES_CASE(4)
    RES_CASE(3) RES_CASE(2) RES_CASE(1) RES_CASE(0)
    default: break;
  }
}
#undef RES_CASE

// ---- TMA staging --------------------------------------------------------------------------------
// One mbarrier; every thread tracks its phase parity.  tma_issue: ONE thread, after a CTA barrier
// that retired all earlier accesses to the destination.
struct Stage {
  uint64_t* bar;
  unsigned par;
  int* status;                  // the evaluation's status word (watchdog)
  long long spin_limit;
};
__device__ __forceinline__ void tma_issue(const Stage& S, double* dst, const double* src, int nblk) {
  asm volatile("fence.proxy.async;\n" ::: "memory");      // generic-proxy writes (this CTA's, or acquired) before the copy
  const uint32_t bytes = (uint32_t)nblk * RBLK * 8;
  mbar_expect_tx(S.bar, bytes);
  bulk_g2s(dst, src, bytes, S.bar);
}
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P1;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void tma_wait(Stage& S) {
  if (!mbar_try(S.bar, S.par & 1u)) {
    const long long t0 = clock64();
    while (!mbar_try(S.bar, S.par & 1u)) {
      if (clock64() - t0 > S.spin_limit) {
        if (S.status) atomicOr(S.status, ST_TIMEOUT);
        break;
      }
    }
  }
  S.par ^= 1u;
}

// Bring rows [0, nrows) of a packed operand (row r = rowlen(r) blocks at block offset rowpos(r),
// rows contiguous) through the staging area `area` (cap blocks) in as few pieces as fit, and run
// body(r0, r1, base) on each piece (base = address of row r0).  The first piece may have been
// issued ahead by the caller (`ahead`, with the same area / cap).  CTA-wide: every thread calls it.
template <class RowLen, class RowPos, class Body>
__device__ __forceinline__ void staged_rows(Stage& S, double* area, int cap, const double* src, int nrows,
                                            RowLen rowlen, RowPos rowpos, bool ahead, Body body) {
  int r = 0;
  while (r < nrows) {
    int r1 = r, nb = 0;
    while (r1 < nrows && nb + rowlen(r1) <= cap) {
      nb += rowlen(r1);
      ++r1;
    }
    if (!ahead) {
      __syncthreads();
      if (threadIdx.x == 0) tma_issue(S, area, src + (long long)rowpos(r) * RBLK, nb);
    }
    ahead = false;
    tma_wait(S);
    body(r, r1, area);
    r = r1;
    if (r < nrows) __syncthreads();
  }
}
// the piece staged_rows() will ask for first: number of blocks of rows [0, r1)
template <class RowLen>
__device__ __forceinline__ int first_piece(int cap, int nrows, RowLen rowlen) {
  int r1 = 0, nb = 0;
  while (r1 < nrows && nb + rowlen(r1) <= cap) {
    nb += rowlen(r1);
    ++r1;
  }
  return nb;
}

// ---- per-unit context (shared memory) -----------------------------------------------------------------
// Every phase is a separate __noinline__ function that reads what it needs from this record: the
// phases then do not compete for registers (as ONE inlined function the per-warp fragment arrays
// were demoted to local memory and the lane offsets recomputed at every use).
struct Ctx {
  int uid, bi, bj, a, b, ab, bb, nr, nyb, pair, is_export, nF, want_grad, dx, dy, r1g;
  int defer;                        // pair whose parent was not ready at the start: W_i is requested in ph_lji, after K_ji
  int oXS, oR1, oR2, oF;            // offsets (doubles) into the dynamic shared memory
  long long ia, ja;
  double* R1g;                      // R1 in this CTA's scratch (units of res_class 1)
  const double* pexp;
  double *oexp, *Zy, *Arow, *Kjj, *Kji, *colp, *taskp, *gx;
};
constexpr int OFF_MISC = 0;
constexpr int OFF_XS = OFF_MISC + R_MISC_DOUBLES;
constexpr int OFF_WB = R_SMEM_BYTES / 8 - R_WB_BLOCKS * RBLK;     // work buffer at the end: F | WB is contiguous
constexpr int WB_TAIL_BLOCKS = 20;                                 // of WB, kept by the gradient phase:
constexpr int OFF_TT = OFF_WB + (R_WB_BLOCKS - WB_TAIL_BLOCKS) * RBLK;   //   task table (4 blocks = 512 ints)
constexpr int OFF_TS = OFF_TT + 4 * RBLK;                                //   one transpose block per warp
static_assert(R_CAP_DOUBLES % RBLK == 0, "region sizes are whole blocks");
constexpr int MISC_Q = 8;             // [RNW] partial |Z|^2
constexpr int MISC_LD = 24;           // [RNW] partial log det
constexpr int MISC_CTX = 40;          // doubles
constexpr int MISC_IDX = 120;         // [2 * EMAXB * 8] ints: global point index or -1
constexpr int MISC_PARAMS = 288;
constexpr int MISC_ROWCNT = 400;      // [2 * EMAXB] ints: gradient tasks per block row
static_assert(sizeof(Ctx) <= (MISC_IDX - MISC_CTX) * 8, "Ctx does not fit");
static_assert(sizeof(ResParams) <= (MISC_ROWCNT - MISC_PARAMS) * 8, "ResParams copy does not fit");
static_assert(MISC_ROWCNT + EMAXB <= R_MISC_DOUBLES, "MISC too small");

__device__ __forceinline__ int* misc_int(int dbl_off, int idx) {
  return reinterpret_cast<int*>(g_smem + OFF_MISC + dbl_off) + idx;
}
#define S_FAIL misc_int(4, 1)
#define S_TCUR misc_int(5, 0)
#define S_TASK misc_int(5, 1)
#define S_NTASK misc_int(6, 0)
#define S_IDX misc_int(MISC_IDX, 0)
#define S_ROWCNT misc_int(MISC_ROWCNT, 0)

constexpr int RTRACE_SLOTS = 512;
// Debug timeline: thread 0 appends (unit << 16 | tag, %globaltimer) to this CTA's slots.
__device__ __forceinline__ void rtrace(const ResParams& P, const Ctx& c, int tag) {
  if (P.trace && threadIdx.x == 0 && *S_TCUR < RTRACE_SLOTS) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    unsigned long long* wp = P.trace + ((long long)blockIdx.x * RTRACE_SLOTS + *S_TCUR) * 2;
    wp[0] = ((unsigned long long)c.uid << 16) | (unsigned long long)tag;
    wp[1] = t;
    ++*S_TCUR;
  }
}

// R1 is addressed either as a shared-memory offset (int) or as a pointer into the scratch (class 1)
struct R1Smem {
  typedef int T;
  static __device__ __forceinline__ int get(const Ctx& c) { return c.oR1; }
};
struct R1Glob {
  typedef double* T;
  static __device__ __forceinline__ double* get(const Ctx& c) { return c.R1g; }
};

static __device__ __noinline__ void dbg_dump(const ResParams& P, const Ctx& c, int phase) {
  if (P.dbg_unit != c.uid || P.dbg_phase != phase) return;
  __syncthreads();
  for (int e = threadIdx.x; e < 160 * 160; e += RNT) {
    const int r = e / 160, cc = e % 160;
    double v1 = 0.0, v2 = 0.0;
    if (c.pair && r < c.bb * 8 && cc < c.ab * 8) {
      const int o = ((r >> 3) * c.ab + (cc >> 3)) * RBLK + sw_off(r & 7, cc & 7);
      v1 = c.r1g ? c.R1g[o] : g_smem[c.oR1 + o];
    }
    if (r < c.bb * 8 && cc < c.bb * 8 && (cc >> 3) <= (r >> 3))
      v2 = g_smem[c.oR2 + (rtri(r >> 3) + (cc >> 3)) * RBLK + sw_off(r & 7, cc & 7)];
    P.dbg_out[e] = v1;
    P.dbg_out[160 * 160 + e] = v2;
  }
  __syncthreads();
}

// ---- 1 x NJ register tile --------------------------------------------------------------------------------
// acc[j] += sum_{k0 <= k < k1} A(k) B_j(k) for j < NJ: one A fragment per k feeds up to four
// independent accumulator chains (the DMMA latency is ~37 cycles, the chains hide it).
//   A(k)   = block at pa + k * sa          (AT: stored [k][m], read transposed)
//   B_j(k) = block at pb[j] + k * sbk      (BT: stored [k][n], read transposed)
// PA / PB: int (shared-memory offset) or pointer (HBM / L2).
template <bool AT, bool BT, int NJ, class PA, class PB>
__device__ __forceinline__ void mk_loop(double2 (&acc)[4], PA pa, int sa, const PB (&pb)[4], int sbk, int k0, int k1,
                                        const Lane& L) {
#pragma unroll 2
  for (int k = k0; k < k1; ++k) {
    const double2 av = AT ? ldt(pa + k * sa, L) : ldn(pa + k * sa, L);
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const double2 bv = BT ? ldt(pb[j] + k * sbk, L) : ldn(pb[j] + k * sbk, L);
      mma2(acc[j], av, bv);
    }
  }
}
// both operands in shared memory: running window addresses, one add per load
template <bool AT, bool BT, int NJ>
__device__ __forceinline__ void mk_loop(double2 (&acc)[4], int pa, int sa, const int (&pb)[4], int sbk, int k0, int k1,
                                        const Lane& L) {
  const unsigned oa = 8u * (unsigned)(pa + k0 * sa);
  unsigned a0 = (AT ? L.bt0 : L.bn) + oa, a1 = L.bt1 + oa;
  unsigned b0[NJ], b1[NJ];
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    const unsigned ob = 8u * (unsigned)(pb[j] + k0 * sbk);
    b0[j] = (BT ? L.bt0 : L.bn) + ob;
    b1[j] = L.bt1 + ob;
  }
  const unsigned sab = 8u * (unsigned)sa, sbb = 8u * (unsigned)sbk;
#pragma unroll 1
  for (int k = k0; k < k1; ++k) {
    double2 av;
    if (AT) {
      av = make_double2(lds64(a0), lds64(a1));
      a1 += sab;
    } else {
      av = lds128(a0);
    }
    a0 += sab;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      double2 bv;
      if (BT) {
        bv = make_double2(lds64(b0[j]), lds64(b1[j]));
        b1[j] += sbb;
      } else {
        bv = lds128(b0[j]);
      }
      b0[j] += sbb;
      mma2(acc[j], av, bv);
    }
  }
}
template <bool AT, bool BT, class PA, class PB>
__device__ __forceinline__ void mk(double2 (&acc)[4], PA pa, int sa, const PB (&pb)[4], int sbk, int k0, int k1, int nj,
                                   const Lane& L) {
  switch (nj) {
    case 4: mk_loop<AT, BT, 4>(acc, pa, sa, pb, sbk, k0, k1, L); break;
    case 3: mk_loop<AT, BT, 3>(acc, pa, sa, pb, sbk, k0, k1, L); break;
    case 2: mk_loop<AT, BT, 2>(acc, pa, sa, pb, sbk, k0, k1, L); break;
    case 1: mk_loop<AT, BT, 1>(acc, pa, sa, pb, sbk, k0, k1, L); break;
    default: break;
  }
}
__device__ __forceinline__ void zero4(double2 (&acc)[4]) {
#pragma unroll
  for (int j = 0; j < 4; ++j) acc[j] = make_double2(0.0, 0.0);
}

// covariance fragment of block (rb, cb) in local block coordinates: rows 8 rb + g, cols 8 cb + 2q, +1
template <int DFN, int WFN>
__device__ __forceinline__ double2 cov_frag(const ResParams& P, const Ctx& c, const Lane& L, int rb, int cb,
                                            bool with_diag) {
  const int tr = rb * 8 + L.g;
  const int tc = cb * 8 + 2 * L.q;
  const int* IDX = S_IDX;
  const double* XS = g_smem + c.oXS;
  const bool rv = IDX[tr] >= 0;
  const double* xr = XS + tr * XD;
  double v0 = cov_value<DFN, WFN>(xr, XS + tc * XD, P.cp);
  double v1 = cov_value<DFN, WFN>(xr, XS + (tc + 1) * XD, P.cp);
  v0 = (rv && IDX[tc] >= 0) ? v0 : 0.0;
  v1 = (rv && IDX[tc + 1] >= 0) ? v1 : 0.0;
  if (with_diag) {
    if (tr == tc) v0 = rv ? P.cp.s2 + P.cp.nv : 1.0;
    if (tr == tc + 1) v1 = rv ? P.cp.s2 + P.cp.nv : 1.0;
  }
  return make_double2(v0, v1);
}

__device__ __forceinline__ int tri_len(int r) { return r + 1; }
// Task of warp w in round i when tasks are ordered by length: rounds alternate direction (snake), so
// that every warp's tasks add up to about the same length (a phase lasts as long as its most loaded warp).
__device__ __forceinline__ int snake_task(int w, int i) { return i * RNW + ((i & 1) ? RNW - 1 - w : w); }
// n columns in ng = ceil(n / 4) groups of nearly equal width (<= 4): first column of group g
__device__ __forceinline__ int grp_lo(int g, int n, int ng) { return (g * n) / ng; }

// ---- P0: gather the unit's coordinate records --------------------------------------------------------------
template <int DFN>
static __device__ __noinline__ void ph_gather(const ResParams& P, const Ctx& c) {
  const int tid = threadIdx.x;
  int* IDX = S_IDX;
  double* XS = g_smem + c.oXS;
  if (tid == 0) *S_FAIL = 0;
  // euclidean distances depend on differences only: coordinates are taken relative to the unit's first
  // point, which keeps the products x_p x_q the gradient phase expands (x_p - x_q) into small
  double ref[MAX_DX] = {0.0, 0.0, 0.0};
  if (DFN == DFN_EUCLIDEAN) {
    const long long i0 = P.perm[c.a > 0 ? c.ia : c.ja];
    for (int d = 0; d < c.dx; ++d) ref[d] = P.X[i0 * c.dx + d];
  }
  for (int t = tid; t < c.nr * 8; t += RNT) {
    long long idx = -1;
    if (t < c.ab * 8) {
      if (t < c.a) idx = P.perm[c.ia + t];
    } else if (t - c.ab * 8 < c.b) {
      idx = P.perm[c.ja + (t - c.ab * 8)];
    }
    IDX[t] = (int)idx;
    double rec[XD];
#pragma unroll
    for (int d = 0; d < MAX_DX + 1; ++d) rec[d] = (idx >= 0 && d < c.dx) ? P.X[idx * c.dx + d] - (d < MAX_DX ? ref[d] : 0.0) : 0.0;
    point_terms(DFN, rec);
#pragma unroll
    for (int d = 0; d < XD; ++d) XS[t * XD + d] = rec[d];
  }
  __syncthreads();
}

// ---- P1: K_ji, L_ji = K_ji W_i^T -------------------------------------------------------------------------------
template <int DFN, int WFN, class R1>
static __device__ __noinline__ void ph_lji(const ResParams& P, const Ctx& c, Stage& stage) {
  const Lane L = make_lane();
  const int w = L.w;
  const int ab = c.ab, bb = c.bb;
  const typename R1::T r1 = R1::get(c);
  // K_ji -> R1 (and to scratch for the gradient contraction): two blocks per step, so that four
  // exponentials are in flight per thread; the stores come after the evaluations (a shared-memory
  // store in between would order the next block's coordinate loads behind it)
  for (int t = w; t < bb * ab; t += 2 * RNW) {
    const int t2 = t + RNW;
    const int row = t / ab, k = t - row * ab;
    const int row2 = t2 / ab, k2 = t2 - row2 * ab;
    const bool has2 = t2 < bb * ab;
    const double2 kv = cov_frag<DFN, WFN>(P, c, L, ab + row, k, false);
    const double2 kv2 = cov_frag<DFN, WFN>(P, c, L, ab + (has2 ? row2 : row), has2 ? k2 : k, false);
    stn(r1 + t * RBLK, L, kv);
    if (c.want_grad) stn(c.Kji + t * RBLK, L, kv);
    if (has2) {
      stn(r1 + t2 * RBLK, L, kv2);
      if (c.want_grad) stn(c.Kji + t2 * RBLK, L, kv2);
    }
  }
  __syncthreads();
  rtrace(P, c, 34);
  if (c.defer && threadIdx.x == 0) {  // the parent was still running when this unit started
    const long long t0 = clock64();
    while (atomicAdd(P.ready0 + c.bi, 0) != P.epoch) {
      __nanosleep(100);
      if (clock64() - t0 > P.spin_limit) {
        atomicOr(P.status, ST_TIMEOUT);
        break;
      }
    }
    __threadfence();
    tma_issue(stage, g_smem + c.oR2, c.pexp + EXP_W, rtri(ab));
  }
  tma_wait(stage);                    // W_i, issued by run_unit (or just now) into [R2 | F]
  rtrace(P, c, 35);
  // L_ji(row, c0 .. c0+3) = sum_{k <= c} K_ji(row, k) W_i(c, k)^T, in place: rounds of whole rows; every
  // task of a round holds its blocks until all of the round's reads are done
  const int Wst = c.oR2;
  const int ng = (ab + 3) >> 2;
  const int rows_round = max(1, (4 * RNW) / ng);
  for (int r0 = 0; r0 < bb; r0 += rows_round) {
    const int ntask = min(rows_round, bb - r0) * ng;
    double2 acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      zero4(acc[i]);
      const int t = w + i * RNW;
      if (t < ntask) {
        const int row = r0 + t / ng, gq = (t % ng + t / ng + (t / ng) / ng) % ng;   // rotated: every warp gets every column group
        const int c0 = grp_lo(gq, ab, ng), nj = grp_lo(gq + 1, ab, ng) - c0;
        const typename R1::T pa = r1 + row * ab * RBLK;
        int pb[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) pb[j] = Wst + rtri(c0 + min(j, nj - 1)) * RBLK;
        mk<false, false>(acc[i], pa, RBLK, pb, RBLK, 0, c0 + 1, nj, L);
#pragma unroll
        for (int j = 1; j < 4; ++j)                    // column c0 + j also takes k = c0 + 1 .. c0 + j
          if (j < nj)
            for (int k = c0 + 1; k <= c0 + j; ++k) mma2(acc[i][j], ldn(pa + k * RBLK, L), ldn(pb[j] + k * RBLK, L));
      }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int t = w + i * RNW;
      if (t < ntask) {
        const int row = r0 + t / ng, gq = (t % ng + t / ng + (t / ng) / ng) % ng;
        const int c0 = grp_lo(gq, ab, ng), nj = grp_lo(gq + 1, ab, ng) - c0;
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (j < nj) stn(r1 + (row * ab + c0 + j) * RBLK, L, acc[i][j]);
      }
    }
    __syncthreads();
  }
}

// ---- P2: S = K_jj + nv I - L_ji L_ji^T (lower blocks, tasks of up to 4 columns) --------------------------
template <int DFN, int WFN, class R1>
static __device__ __noinline__ void ph_schur(const ResParams& P, const Ctx& c) {
  const Lane L = make_lane();
  const int ab = c.ab, bb = c.bb;
  const typename R1::T r1 = R1::get(c);
  // task list: rows descending (longest first), column groups of nearly equal width (<= 4)
  rtrace(P, c, 45);
  int t = 0;
  for (int row = bb - 1; row >= 0; --row) {
    const int ng = (row + 4) >> 2;
    for (int gq = 0; gq < ng; ++gq, ++t) {
      if ((t & (RNW - 1)) != L.w) continue;
      const int c0 = grp_lo(gq, row + 1, ng), nj = grp_lo(gq + 1, row + 1, ng) - c0;
      // the covariance values first (independent chains, nothing stored in between), then the product
      double2 kv[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) kv[j] = cov_frag<DFN, WFN>(P, c, L, ab + row, ab + c0 + min(j, nj - 1), true);
      double2 acc[4];
      zero4(acc);
      if (c.pair) {
        typename R1::T pb[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) pb[j] = r1 + (c0 + min(j, nj - 1)) * ab * RBLK;
        mk<false, false>(acc, r1 + row * ab * RBLK, RBLK, pb, RBLK, 0, ab, nj, L);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (j < nj) {
          if (c.want_grad) stn(c.Kjj + (rtri(row) + c0 + j) * RBLK, L, kv[j]);
          stn(c.oR2 + (rtri(row) + c0 + j) * RBLK, L, make_double2(kv[j].x - acc[j].x, kv[j].y - acc[j].y));
        }
      }
    }
  }
  rtrace(P, c, 46);
  __syncthreads();
}

// ---- P3: Cholesky of S and W_S = L_S^-1, both in place -----------------------------------------------------
// Named barrier 1: warp 0 arrives, warps 1.. wait (producer / consumer; barrier 0 is __syncthreads).
__device__ __forceinline__ void nbar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void nbar_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// Diagonal block J by ONE warp: L_JJ in place, W_JJ = L_JJ^-1 to WD.  Every lane holds the whole block.
__device__ __forceinline__ void chol_diag_block(int blk, int wd, int row0, bool store_all) {
  const int lane = threadIdx.x & 31;
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
  __syncwarp();                                  // every lane has read the block before any lane writes into it
  const int f = chol8_full(A, wv, lane);
  if (lane == 0 && f != 0 && *S_FAIL == 0) *S_FAIL = row0 + f;
  if (lane < 8) {
#pragma unroll
    for (int v = 0; v < 8; ++v) g_smem[wd + sw_off(v, lane)] = wv[v];   // lane c holds column c of W_JJ
  }
  // Of L_JJ only the diagonal is used afterwards (log-determinant); the whole block is written for
  // the debug dump.  Every lane holds the same L: same-value stores, upper part zeroed.
  if (!store_all) {
#pragma unroll
    for (int r = 0; r < 8; ++r) g_smem[blk + sw_off(r, r)] = A[r][r];      // (a lane == r test compiles to a jump table)
    return;
  }
#pragma unroll
  for (int r = 0; r < 8; ++r)
#pragma unroll
    for (int cp = 0; cp < 4; ++cp) {
      const double v0 = (2 * cp <= r) ? A[r][2 * cp] : 0.0;
      const double v1 = (2 * cp + 1 <= r) ? A[r][2 * cp + 1] : 0.0;
      *reinterpret_cast<double2*>(g_smem + blk + sw_off(r, 2 * cp)) = make_double2(v0, v1);
    }
}

// Right-looking factorisation with the diagonal blocks looked ahead: in step J warp 0 alone carries
// the chain  L_{J+1,J} -> C_{J+1,J+1} -> chol / inverse of diagonal block J+1  while the other warps
// do the rest of panel J and of the trailing update (they form L_{J+1,J} in registers themselves), so
// that a step costs one diagonal block (~1 us) and ONE CTA barrier.
static __device__ __noinline__ void ph_chol_inv(const ResParams& P, const Ctx& c) {
  const Lane L = make_lane();
  const int tid = threadIdx.x, w = L.w;
  const int bb = c.bb;
  const int R2 = c.oR2;
  const int WD = OFF_WB;                           // diagonal-block inverses, in WB
  const bool store_all = P.dbg_unit == c.uid;
  // Warp 0's scalar FP64 chain shares its sub-partition's FP64 pipe with the DMMAs of warps 4, 8, 12
  // (measured: the diagonal block takes 2180 cycles next to them, 1250 alone): those three sit out.
  const int wk = (w & 3) ? w - 1 - (w >> 2) : -1;       // worker index 0 .. NWK-1 of warps 1,2,3,5,6,7,...
  constexpr int NWK = RNW - RNW / 4;
  if (w == 0) chol_diag_block(R2, WD, 0, store_all);
  for (int J = 0; J + 1 < bb; ++J) {
    __syncthreads();                               // L_JJ, W_JJ and the trailing update of step J - 1
    rtrace(P, c, 30);
    if (wk < 0 && w != 0) {
      nbar_arrive(1, RNT);
      continue;
    }
    const int p10 = R2 + (rtri(J + 1) + J) * RBLK;
    const double2 bw = ldn(WD + J * RBLK, L);
    const double2 c10 = ldn(p10, L);
    double2 cv = make_double2(0.0, 0.0);
    if (w == 0) cv = ldn(p10 + RBLK, L);           // C_{J+1,J+1}: complete since the barrier
    double2 lp = make_double2(0.0, 0.0);
    mma2(lp, c10, bw);                             // L_{J+1,J}: every warp forms it in registers
    if (w == 0) {
      nbar_arrive(1, RNT);                         // C_{J+1,J} has been read by this warp
      const int p11 = p10 + RBLK;
      mma2(cv, neg2(lp), lp);
      stn(p11, L, cv);
      __syncwarp();
      rtrace(P, c, 31);
      chol_diag_block(p11, WD + (J + 1) * RBLK, (J + 1) * 8, store_all);
      rtrace(P, c, 32);
    } else {
      // panel: L_IJ = C_IJ W_JJ^T, rows I >= J + 2
      for (int I = J + 2 + wk; I < bb; I += NWK) {
        const int pc = R2 + (rtri(I) + J) * RBLK;
        double2 o = make_double2(0.0, 0.0);
        mma2(o, ldn(pc, L), bw);
        stn(pc, L, o);
      }
      nbar_sync(1, RNT);                           // panel J complete; every warp holds L_{J+1,J}
      if (wk == 0) stn(p10, L, lp);
      // trailing update: C_IK -= L_IJ L_KJ^T for rows I >= J + 2, columns J + 1 <= K <= I
      const int m = bb - J - 2;
      const int ntask = m * (m + 3) / 2;
      for (int t = wk; t < ntask; t += NWK) {
        int ii = (int)((sqrtf(8.0f * (float)t + 9.0f) - 3.0f) * 0.5f);
        while ((ii + 1) * (ii + 4) / 2 <= t) ++ii;
        while (ii * (ii + 3) / 2 > t) --ii;
        const int kk = t - ii * (ii + 3) / 2;
        const int I = J + 2 + ii, K = J + 1 + kk;
        const int pc = R2 + (rtri(I) + K) * RBLK;
        double2 cu = ldn(pc, L);
        const double2 bv = (kk == 0) ? lp : ldn(R2 + (rtri(K) + J) * RBLK, L);
        mma2(cu, neg2(ldn(R2 + (rtri(I) + J) * RBLK, L)), bv);
        stn(pc, L, cu);
      }
    }
  }
  __syncthreads();
  dbg_dump(P, c, 3);
  rtrace(P, c, 4);
  // log-determinant: sum of log L_tt over the unit's j rows, fixed order
  {
    double lv = 0.0;
    if (tid < bb * 8) lv = log(g_smem[R2 + (rtri(tid >> 3) + (tid >> 3)) * RBLK + sw_off(tid & 7, tid & 7)]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) lv += __shfl_xor_sync(0xffffffffu, lv, o);
    if (L.lane == 0) g_smem[OFF_MISC + MISC_LD + w] = lv;
  }
  // W_S = L_S^-1 in place, by descending block columns.  Row I belongs to warp I mod RNW for the whole
  // phase: the W blocks a task reads are its own earlier results, the L blocks it reads (column K) are
  // overwritten only after the step's barrier - one barrier per step.
  for (int K = bb - 1; K >= 0; --K) {
    double2 o[2];
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      o[rr] = make_double2(0.0, 0.0);
      const int I = w + rr * RNW;
      if (I > K && I < bb) {
        double2 a[4];
        zero4(a);
        int J = K + 1;
        for (; J + 3 <= I; J += 4) {
#pragma unroll
          for (int u = 0; u < 4; ++u) mma2(a[u], ldn(R2 + (rtri(I) + J + u) * RBLK, L), ldt(R2 + (rtri(J + u) + K) * RBLK, L));
        }
#pragma unroll
        for (int u = 0; u < 3; ++u)
          if (J + u <= I) mma2(a[u], ldn(R2 + (rtri(I) + J + u) * RBLK, L), ldt(R2 + (rtri(J + u) + K) * RBLK, L));
        a[0].x = (a[0].x + a[1].x) + (a[2].x + a[3].x);
        a[0].y = (a[0].y + a[1].y) + (a[2].y + a[3].y);
        mma2(o[rr], a[0], ldt(WD + K * RBLK, L));
      }
    }
    __syncthreads();
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      const int I = w + rr * RNW;
      if (I > K && I < bb) stn(R2 + (rtri(I) + K) * RBLK, L, neg2(o[rr]));
    }
    if ((K & (RNW - 1)) == w) stn(R2 + (rtri(K) + K) * RBLK, L, ldn(WD + K * RBLK, L));
  }
  __syncthreads();
  if (c.is_export) {                                // W_b for this block's pairs
    double* dst = c.oexp + EXP_W;
    const double* src = g_smem + R2;
    for (int e = tid; e < rtri(bb) * RBLK / 2; e += RNT)
      reinterpret_cast<double2*>(dst)[e] = reinterpret_cast<const double2*>(src)[e];
    if (P.ready0) {
      // The pairs of this block can start now: L_ji, S and the factorisation of S need W_i only; Z_i, the
      // alpha rows and the scalars follow under `ready` (the pair waits for it before it requests Z_i).
      __threadfence();
      __syncthreads();
      if (tid == 0) atomicExch(P.ready0 + c.bj, P.epoch);
    }
  }
}

// ---- P4: Z_j = W_S (Y_j - L_ji Z_i), alpha_j = W_S^T Z_j, the log-likelihood ---------------------------------
// Pieces of up to nyc blocks of 8 outputs (normally all of them at once).  Z_i of the piece is staged in F;
// the right-hand sides then take its place, as (yl * bb + row).  Tasks = (row, 2 output blocks), dealt
// round-robin; the in-place steps hold their results in registers until every read is done.
constexpr int YSLOTS = 5;             // tasks per warp and piece: bb * ceil(ny / 2) <= 20 * 4 = 80 <= 16 * 5
template <class R1>
static __device__ __noinline__ void ph_ypart(const ResParams& P, const Ctx& c, Stage& stage, int nyc) {
  const Lane L = make_lane();
  const int tid = threadIdx.x, w = L.w;
  const int ab = c.ab, bb = c.bb, nyb = c.nyb;
  const typename R1::T r1 = R1::get(c);
  const int RB = c.oF, R2 = c.oR2;
  const int* IDX = S_IDX;
  double qsum = 0.0;
  bool ahead = c.pair != 0;
  for (int y0 = 0; y0 < nyb; y0 += nyc) {
    const int ny = min(nyc, nyb - y0);
    const int ngy = (ny + 1) >> 1;
    const int ntask = bb * ngy;
    // R = Y_j - L_ji Z_i.  The Y values of all the warp's tasks are requested first: they come from
    // HBM / L2 and would otherwise be waited for task by task.
    double2 z[YSLOTS][2];
#pragma unroll
    for (int i = 0; i < YSLOTS; ++i) {
      const int t = snake_task(w, i);
      z[i][0] = z[i][1] = make_double2(0.0, 0.0);
      if (t < ntask) {
        const int row = t / ngy, yl0 = ((t + row + row / ngy) % ngy) * 2;     // output groups rotated by row
        const int nj = min(2, ny - yl0);
        const int idx = IDX[ab * 8 + row * 8 + L.g];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int yc = (y0 + yl0 + j) * 8 + 2 * L.q;
          if (idx >= 0 && j < nj) {
            if (yc < c.dy) z[i][j].x = __ldg(P.Y + (long long)idx * c.dy + yc);
            if (yc + 1 < c.dy) z[i][j].y = __ldg(P.Y + (long long)idx * c.dy + yc + 1);
          }
        }
      }
    }
    if (c.pair) {
      if (!ahead && tid == 0) tma_issue(stage, g_smem + c.oF, c.pexp + EXP_ZY + (long long)y0 * ab * RBLK, ny * ab);
      ahead = false;
      tma_wait(stage);
#pragma unroll
      for (int i = 0; i < YSLOTS; ++i) {
        const int t = snake_task(w, i);
        if (t < ntask) {
          const int row = t / ngy, yl0 = ((t + row + row / ngy) % ngy) * 2;     // output groups rotated by row
          const int nj = min(2, ny - yl0);
          double2 acc[4];
          zero4(acc);
          int pb[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) pb[j] = c.oF + (yl0 + min(j, nj - 1)) * ab * RBLK;
          mk<false, true>(acc, r1 + row * ab * RBLK, RBLK, pb, RBLK, 0, ab, nj, L);
          z[i][0].x -= acc[0].x;
          z[i][0].y -= acc[0].y;
          z[i][1].x -= acc[1].x;
          z[i][1].y -= acc[1].y;
        }
      }
    }
    __syncthreads();                                 // Z_i is no longer needed: the right-hand sides replace it
    rtrace(P, c, 36);
#pragma unroll
    for (int i = 0; i < YSLOTS; ++i) {
      const int t = snake_task(w, i);
      if (t < ntask) {
        const int row = t / ngy, yl0 = ((t + row + row / ngy) % ngy) * 2;     // output groups rotated by row
#pragma unroll
        for (int j = 0; j < 2; ++j)
          if (yl0 + j < ny) stn(RB + ((yl0 + j) * bb + row) * RBLK, L, z[i][j]);
      }
    }
    __syncthreads();
    // Z = W_S R  (held in registers until every read of R is done)
#pragma unroll
    for (int i = 0; i < YSLOTS; ++i) {
      const int t = snake_task(w, i);
      if (t < ntask) {
        const int row = t / ngy, yl0 = ((t + row + row / ngy) % ngy) * 2;     // output groups rotated by row
        const int nj = min(2, ny - yl0);
        double2 acc[4];
        zero4(acc);
        int pb[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) pb[j] = RB + (yl0 + min(j, nj - 1)) * bb * RBLK;
        mk<false, true>(acc, R2 + rtri(row) * RBLK, RBLK, pb, RBLK, 0, row + 1, nj, L);
        z[i][0] = acc[0];
        z[i][1] = acc[1];
        qsum += acc[0].x * acc[0].x + acc[0].y * acc[0].y;
        if (nj > 1) qsum += acc[1].x * acc[1].x + acc[1].y * acc[1].y;
      }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < YSLOTS; ++i) {
      const int t = snake_task(w, i);
      if (t < ntask) {
        const int row = t / ngy, yl0 = ((t + row + row / ngy) % ngy) * 2;     // output groups rotated by row
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          if (yl0 + j < ny) {
            stn(RB + ((yl0 + j) * bb + row) * RBLK, L, z[i][j]);
            stn(c.Zy + ((y0 + yl0 + j) * bb + row) * RBLK, L, z[i][j]);
          }
        }
      }
    }
    __syncthreads();
    rtrace(P, c, 37);
    // alpha_j = W_S^T Z
    if (c.want_grad) {
      for (int i = 0; i * RNW < ntask; ++i) {
        const int t = snake_task(w, i);
        if (t >= ntask) continue;
        const int row = t / ngy, yl0 = ((t + row + row / ngy) % ngy) * 2;     // output groups rotated by row
        const int nj = min(2, ny - yl0);
        double2 a0 = make_double2(0.0, 0.0), a1 = make_double2(0.0, 0.0);
        const int p0 = RB + yl0 * bb * RBLK, p1 = RB + (yl0 + nj - 1) * bb * RBLK;
        // A(k) = W_S(k, row) read transposed: block (k, row) of the packed lower triangle
#pragma unroll 2
        for (int k = row; k < bb; ++k) {
          const double2 av = ldt(R2 + (rtri(k) + row) * RBLK, L);
          mma2(a0, av, ldt(p0 + k * RBLK, L));
          mma2(a1, av, ldt(p1 + k * RBLK, L));
        }
        stn(c.Arow + ((long long)(ab + row) * nyb + y0 + yl0) * RBLK, L, a0);
        if (nj > 1) stn(c.Arow + ((long long)(ab + row) * nyb + y0 + yl0 + 1) * RBLK, L, a1);
      }
    }
    __syncthreads();
  }
  // |Z_j|^2 and the log-likelihood (gprf.py:542-544)
  double qv = qsum;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) qv += __shfl_xor_sync(0xffffffffu, qv, o);
  if (L.lane == 0) g_smem[OFF_MISC + MISC_Q + w] = qv;
  __syncthreads();
  if (tid == 0) {
    double qt = 0.0, lt = 0.0;
    for (int i = 0; i < RNW; ++i) {
      qt += g_smem[OFF_MISC + MISC_Q + i];
      lt += g_smem[OFF_MISC + MISC_LD + i];
    }
    double logdet = 2.0 * lt;
    if (c.pair) {
      qt += c.pexp[EXP_SCAL + 1];
      logdet += c.pexp[EXP_SCAL + 0];
    }
    if (c.is_export) {
      c.oexp[EXP_SCAL + 0] = logdet;
      c.oexp[EXP_SCAL + 1] = qt;
    }
    P.ll_u[c.uid] = -0.5 * qt - 0.5 * c.dy * logdet - 0.5 * c.dy * (double)(c.a + c.b) * 1.8378770664093454836;
    if (*S_FAIL != 0) {
      P.info[c.uid] = *S_FAIL;
      atomicOr(P.status, ST_NOTPD);
    }
  }
}

// ---- P6a: T = L_ji W_i, in place (tasks of 4 columns, rounds of whole rows, W_i rows staged in F) -----------
template <class R1>
static __device__ __noinline__ void ph_t(const ResParams& P, const Ctx& c, Stage& stage) {
  const Lane L = make_lane();
  const int w = L.w;
  const int ab = c.ab, bb = c.bb, nF = c.nF;
  const typename R1::T r1 = R1::get(c);
  const int ng = (ab + 3) >> 2;
  const int rows_round = max(1, (4 * RNW) / ng);
  const int oF = c.oF;
  bool ahead = true;
  for (int r0 = 0; r0 < bb; r0 += rows_round) {
    const int ntask = min(rows_round, bb - r0) * ng;
    double2 acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) zero4(acc[i]);
    staged_rows(stage, g_smem + oF, nF, c.pexp + EXP_W, ab, tri_len, rtri, ahead, [&](int k0, int k1, const double*) {
      rtrace(P, c, 42);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int t = w + i * RNW;
        if (t < ntask) {
          const int row = r0 + t / ng, gq = (t % ng + t / ng + (t / ng) / ng) % ng;   // rotated: every warp gets every column group
          const int c0 = grp_lo(gq, ab, ng), nj = grp_lo(gq + 1, ab, ng) - c0;
          const typename R1::T pa = r1 + row * ab * RBLK;
          // T(row, c0 + j) += sum_{k >= c0 + j} L_ji(row, k) W_i(k, c0 + j)
          int k = max(k0, c0);
          for (; k < min(k1, c0 + 3); ++k) {        // rows of W_i that do not reach all four columns yet
            const double2 av = ldn(pa + k * RBLK, L);
            const int pk = oF + (rtri(k) - rtri(k0) + c0) * RBLK;
            const int nk = min(nj, k - c0 + 1);
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (j < nk) mma2(acc[i][j], av, ldt(pk + j * RBLK, L));
          }
          if (nj == 4) {
#pragma unroll 2
            for (; k < k1; ++k) {
              const double2 av = ldn(pa + k * RBLK, L);
              const int pk = oF + (rtri(k) - rtri(k0) + c0) * RBLK;
#pragma unroll
              for (int j = 0; j < 4; ++j) mma2(acc[i][j], av, ldt(pk + j * RBLK, L));
            }
          } else {
            for (; k < k1; ++k) {
              const double2 av = ldn(pa + k * RBLK, L);
              const int pk = oF + (rtri(k) - rtri(k0) + c0) * RBLK;
#pragma unroll
              for (int j = 0; j < 4; ++j)
                if (j < nj) mma2(acc[i][j], av, ldt(pk + j * RBLK, L));
            }
          }
        }
      }
    });
    ahead = false;
    rtrace(P, c, 43);
    __syncthreads();
    rtrace(P, c, 44);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int t = w + i * RNW;
      if (t < ntask) {
        const int row = r0 + t / ng, gq = (t % ng + t / ng + (t / ng) / ng) % ng;
        const int c0 = grp_lo(gq, ab, ng), nj = grp_lo(gq + 1, ab, ng) - c0;
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (j < nj) stn(r1 + (row * ab + c0 + j) * RBLK, L, acc[i][j]);
      }
    }
    __syncthreads();
  }
}

// ---- P6b: V = -W_S T, in place: rounds of whole rows, highest rows first (row I is read by rows >= I only) ---
template <class R1>
static __device__ __noinline__ void ph_v(const ResParams& P, const Ctx& c) {
  const Lane L = make_lane();
  const int w = L.w;
  const int ab = c.ab, bb = c.bb;
  const typename R1::T r1 = R1::get(c);
  const int ng = (ab + 3) >> 2;
  const int rows_round = max(1, (4 * RNW) / ng);
  for (int rhi = bb; rhi > 0; rhi -= rows_round) {
    const int rlo = max(0, rhi - rows_round);
    const int ntask = (rhi - rlo) * ng;
    double2 acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      zero4(acc[i]);
      const int t = snake_task(w, i);
      if (t < ntask) {
        const int row = rhi - 1 - t / ng, gq = (t % ng + t / ng + (t / ng) / ng) % ng;     // longest rows first, groups rotated
        const int c0 = grp_lo(gq, ab, ng), nj = grp_lo(gq + 1, ab, ng) - c0;
        typename R1::T pb[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) pb[j] = r1 + (c0 + min(j, nj - 1)) * RBLK;
        mk<false, true>(acc[i], c.oR2 + rtri(row) * RBLK, RBLK, pb, ab * RBLK, 0, row + 1, nj, L);
      }
    }
    rtrace(P, c, 40);
    __syncthreads();
    rtrace(P, c, 41);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int t = snake_task(w, i);
      if (t < ntask) {
        const int row = rhi - 1 - t / ng, gq = (t % ng + t / ng + (t / ng) / ng) % ng;
        const int c0 = grp_lo(gq, ab, ng), nj = grp_lo(gq + 1, ab, ng) - c0;
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (j < nj) stn(r1 + (row * ab + c0 + j) * RBLK, L, neg2(acc[i][j]));
      }
    }
    __syncthreads();
  }
}

// ---- P7: alpha_i = alpha_i(block) + V^T Z_j  (tasks (row of i, 4 output blocks); Z_j rows staged in F) --------
template <class R1>
static __device__ __noinline__ void ph_alpha_i(const ResParams& P, const Ctx& c, Stage& stage) {
  const Lane L = make_lane();
  const int w = L.w;
  const int ab = c.ab, bb = c.bb, nyb = c.nyb;
  const typename R1::T r1 = R1::get(c);
  const int oF = c.oF;
  auto rl = [&](int) { return bb; };
  auto rp = [&](int r) { return r * bb; };
  staged_rows(stage, g_smem + oF, c.nF, c.Zy, nyb, rl, rp, true, [&](int y0, int y1, const double*) {
    const int ngy = (y1 - y0 + 1) >> 1;
    const int ntask = ab * ngy;
    for (int t = w; t < ntask; t += RNW) {
      const int row = t / ngy, yl0 = (t - row * ngy) * 2;
      const int nj = min(2, y1 - y0 - yl0);
      double2 acc[4];
      zero4(acc);
#pragma unroll
      for (int j = 0; j < 2; ++j)
        if (j < nj) acc[j] = ldn(c.pexp + EXP_AROW + ((long long)row * nyb + y0 + yl0 + j) * RBLK, L);
      int pb[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) pb[j] = oF + (yl0 + min(j, nj - 1)) * bb * RBLK;
      mk<true, true>(acc, r1 + row * RBLK, ab * RBLK, pb, RBLK, 0, bb, nj, L);
#pragma unroll
      for (int j = 0; j < 2; ++j)
        if (j < nj) stn(c.Arow + ((long long)row * nyb + y0 + yl0 + j) * RBLK, L, acc[j]);
    }
  });
}

// ---- P8: G = alpha alpha^T - dy K^-1, contracted with dK in registers (gprf.py:547-584) ------------------
// Column pieces: alpha rows [c0, c1) staged in F | WB (B operands).  Tasks = (block row r, up to GCOLS
// columns of the piece, never straddling the i / j boundary) taken from a shared counter, biggest rows
// first; alpha(r), the saved covariance values and K_ii^-1 come from L2 and are requested before the
// K^-1 products so that their latency hides behind them.
//
// Contraction with dK.  Euclidean distances: with M = G o w (w = k for SE, k / (1 + sqrt3 r) for
// Matern-3/2; strictly lower entries only) every sum over a block is linear in M,
//     sum_q M_pq (x_p - x_q)_d   = x_pd (M 1)_p - (M X)_pd
//     sum_q M_pq (x_p - x_q)_d^2 = x_pd^2 (M 1)_p - 2 x_pd (M X)_pd + (M X^2)_pd
// so one DMMA pair per block, M (already in the A-fragment layout) times the feature block
// FE = [x0 x1 x2 1 x0^2 x1^2 x2^2 0] of the column points, accumulates everything the ROW points need over
// the task's columns, and a second pair, M^T (through a per-warp transpose block) times FE of the row
// points, gives what the COLUMN points need - no shuffles, no per-entry derivative arithmetic.  The
// lengthscale factors are applied by ph_finalize.  (Coordinates are centred, ph_gather.)
// Great-circle distances (lld) keep the per-entry derivative evaluation.
template <int DFN, int WFN, class R1>
static __device__ __noinline__ void ph_grad(const ResParams& P, const Ctx& c, Stage& stage) {
  const Lane L = make_lane();
  const int tid = threadIdx.x;
  const int ab = c.ab, bb = c.bb, nr = c.nr, nyb = c.nyb;
  const typename R1::T r1 = R1::get(c);
  const int R2 = c.oR2, oF = c.oF;
  const double ndy = -(double)c.dy;
  const CovParams& cp = P.cp;
  const int* IDX = S_IDX;
  const double* XS = g_smem + c.oXS;
  const int FE = c.oXS;                            // euclidean: feature blocks take the place of XS
  int* TT = reinterpret_cast<int*>(g_smem + OFF_TT);
  const int TS = OFF_TS + L.w * RBLK;
  int* rowcnt = S_ROWCNT;
  if (tid < nr) rowcnt[tid] = 0;
  if (DFN == DFN_EUCLIDEAN) {
    // FE block of 8 points: element (point, feature), stored like every other block
    double f[8];
    const bool mine = tid < nr * 8;
    if (mine) {
      const double* x = XS + tid * XD;
      f[0] = x[0]; f[1] = x[1]; f[2] = x[2]; f[3] = 1.0;
      f[4] = x[0] * x[0]; f[5] = x[1] * x[1]; f[6] = x[2] * x[2]; f[7] = 0.0;
    }
    __syncthreads();
    if (mine) {
      double* blk = g_smem + FE + (tid >> 3) * RBLK;
#pragma unroll
      for (int n = 0; n < 8; ++n) blk[sw_off(tid & 7, n)] = f[n];
    }
  }
  const int cap = c.nF + R_WB_BLOCKS - WB_TAIL_BLOCKS;
  const int crow = max(GCOLS, (cap / nyb) & ~(GCOLS - 1));     // alpha rows per piece (a multiple of GCOLS)
  for (int c0 = 0; c0 < nr; c0 += crow) {
    const int c1 = min(nr, c0 + crow);
    __syncthreads();
    rtrace(P, c, 19);
    if (tid == 0) tma_issue(stage, g_smem + oF, c.Arow + (long long)c0 * nyb * RBLK, (c1 - c0) * nyb);
    if (L.w == 0) {
      // task table, rows descending: lanes count their rows' groups, prefix sum, then fill
      int cnt[2], r_[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int r = nr - 1 - (L.lane + 32 * h);
        r_[h] = r;
        cnt[h] = 0;
        if (r >= c0) {
          const int ce = min(c1, r + 1);
          const int ni = max(0, min(ce, ab) - c0), nj2 = max(0, ce - max(c0, ab));
          cnt[h] = (ni + GCOLS - 1) / GCOLS + (nj2 + GCOLS - 1) / GCOLS;
        }
      }
      int pre = cnt[0];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, pre, o);
        if (L.lane >= o) pre += v;
      }
      const int tot0 = __shfl_sync(0xffffffffu, pre, 31);
      int pre1 = cnt[1];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, pre1, o);
        if (L.lane >= o) pre1 += v;
      }
      const int tot1 = __shfl_sync(0xffffffffu, pre1, 31);
      int at[2] = {pre - cnt[0], tot0 + pre1 - cnt[1]};
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int r = r_[h];
        if (r >= c0 && cnt[h] > 0) {
          const int ce = min(c1, r + 1);
          int gi = rowcnt[r], nt = at[h];
          for (int part = 0; part < 2; ++part) {     // columns of the i part, then of the j part
            const int lo = part ? max(c0, ab) : c0, hi = part ? ce : min(ce, ab);
            for (int cl = lo; cl < hi; cl += GCOLS, ++gi)
              TT[nt++] = r | (cl << 6) | (min(hi, cl + GCOLS) << 12) | (gi << 18);
          }
          rowcnt[r] = gi;
        }
      }
      if (L.lane == 0) {
        *S_NTASK = tot0 + tot1;
        *S_TASK = 0;
      }
    }
    __syncthreads();
    rtrace(P, c, 20);
    tma_wait(stage);
    rtrace(P, c, 21);
    const int ntask = *S_NTASK;
    while (true) {
      int t = 0;
      if (L.lane == 0) t = atomicAdd(S_TASK, 1);
      t = __shfl_sync(0xffffffffu, t, 0);
      if (t >= ntask) break;
      const int code = TT[t];
      const int r = code & 63, cl = (code >> 6) & 63, ch = (code >> 12) & 63, gi = code >> 18;
      const int nj = ch - cl;
      const bool irow = r < ab;
      const bool icol = cl < ab;
      const int wi = irow ? r : r - ab;            // row inside its part
      // requests to L2 first
      double2 af[RNYB], kv[4], ki[4];
#pragma unroll
      for (int y = 0; y < RNYB; ++y)
        af[y] = (y < nyb) ? ldn(c.Arow + ((long long)r * nyb + y) * RBLK, L) : make_double2(0.0, 0.0);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int cc = cl + min(j, nj - 1);
        ki[j] = make_double2(0.0, 0.0);
        if (irow) {
          ki[j] = ldn(c.pexp + EXP_KINV + (long long)(rtri(wi) + cc) * RBLK, L);
          kv[j] = ldn(c.pexp + EXP_KSAVE + (long long)(rtri(wi) + cc) * RBLK, L);
        } else if (icol) {
          kv[j] = ldn(c.Kji + (wi * ab + cc) * RBLK, L);
        } else {
          kv[j] = ldn(c.Kjj + (rtri(wi) + cc - ab) * RBLK, L);
        }
      }
      // ---- K^-1 blocks (r, cl .. ch-1)
      double2 acc[4];
      zero4(acc);
      if (irow) {                                  // (i row, i columns): V^T V   (+ K_ii^-1)
        typename R1::T pb[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) pb[j] = r1 + (cl + min(j, nj - 1)) * RBLK;
        mk<true, true>(acc, r1 + wi * RBLK, ab * RBLK, pb, ab * RBLK, 0, bb, nj, L);
      } else if (icol) {                           // (j row, i columns): W_S^T V
        typename R1::T pb[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) pb[j] = r1 + (cl + min(j, nj - 1)) * RBLK;
        for (int k = wi; k < bb; ++k) {
          const double2 av = ldt(R2 + (rtri(k) + wi) * RBLK, L);
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (j < nj) mma2(acc[j], av, ldt(pb[j] + k * ab * RBLK, L));
        }
      } else {                                     // (j row, j columns): W_S^T W_S
        for (int k = wi; k < bb; ++k) {
          const int rowk = R2 + rtri(k) * RBLK;
          const double2 av = ldt(rowk + wi * RBLK, L);
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (j < nj) mma2(acc[j], av, ldt(rowk + (cl - ab + j) * RBLK, L));
        }
      }
      const int tr = r * 8 + L.g;
      const bool rv = IDX[tr] >= 0;
      double* tp = c.taskp + ((long long)r * MAXG + gi) * TASKP;
      double s1 = 0.0, d0 = 0.0;                   // sum G k (strictly lower), sum G_pp / 2
      if (DFN == DFN_EUCLIDEAN) {
        double2 orow = make_double2(0.0, 0.0);
        const double2 fr = ldt(FE + r * RBLK, L);  // features of the row points, as a [k][n] operand
        double xr[3];
        if (WFN != WFN_SE) {
#pragma unroll
          for (int d = 0; d < 3; ++d) xr[d] = g_smem[FE + r * RBLK + sw_off(L.g, d)];
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (j >= nj) break;
          const int cc = cl + j;
          double2 g = make_double2(acc[j].x + ki[j].x, acc[j].y + ki[j].y);
          if (!irow && !icol && c.is_export) stn(c.oexp + EXP_KINV + (long long)(rtri(wi) + cc - ab) * RBLK, L, g);
          g.x *= ndy;
          g.y *= ndy;
          double2 g2 = make_double2(0.0, 0.0);
          const int arow = oF + (cc - c0) * nyb * RBLK;
#pragma unroll
          for (int y = 0; y < RNYB; ++y)
            if (y < nyb) mma2((y & 1) ? g2 : g, af[y], ldn(arow + y * RBLK, L));
          g.x += g2.x;
          g.y += g2.y;
          const int tc = cc * 8 + 2 * L.q;
          const bool o0 = rv && IDX[tc] >= 0 && tc < tr, o1 = rv && IDX[tc + 1] >= 0 && tc + 1 < tr;
          if (rv && tc == tr) d0 += 0.5 * g.x;
          if (rv && tc + 1 == tr) d0 += 0.5 * g.y;
          double2 m = make_double2(o0 ? g.x * kv[j].x : 0.0, o1 ? g.y * kv[j].y : 0.0);
          s1 += m.x + m.y;
          if (WFN != WFN_SE) {                     // Matern-3/2: w'(r)/r = -3 k / (1 + sqrt3 r)
            double r2a = 0.0, r2b = 0.0;
#pragma unroll
            for (int d = 0; d < 3; ++d) {
              const double da = xr[d] - g_smem[FE + cc * RBLK + sw_off(2 * L.q, d)];
              const double db = xr[d] - g_smem[FE + cc * RBLK + sw_off(2 * L.q + 1, d)];
              r2a += da * da * cp.il2[d];
              r2b += db * db * cp.il2[d];
            }
            m.x /= 1.0 + GPRF_SQRT3 * sqrt(r2a);
            m.y /= 1.0 + GPRF_SQRT3 * sqrt(r2b);
          }
          mma2(orow, m, ldt(FE + cc * RBLK, L));   // what the row points need, summed over the columns
          __syncwarp();
          stn(TS, L, m);
          __syncwarp();
          double2 ocol = make_double2(0.0, 0.0);
          mma2(ocol, ldt(TS, L), fr);              // what the column points need: (column, feature 2q / 2q+1)
          if (L.q < 2) *reinterpret_cast<double2*>(c.colp + (long long)(rtri(r) + cc) * COLP + L.g * 4 + 2 * L.q) = ocol;
        }
        *reinterpret_cast<double2*>(tp + L.g * 8 + 2 * L.q) = orow;
      } else {
        const double* xr = XS + tr * XD;
        double rs[3] = {0.0, 0.0, 0.0}, thl[3] = {0.0, 0.0, 0.0};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (j >= nj) break;
          const int cc = cl + j;
          double2 g = make_double2(acc[j].x + ki[j].x, acc[j].y + ki[j].y);
          if (!irow && !icol && c.is_export) stn(c.oexp + EXP_KINV + (long long)(rtri(wi) + cc - ab) * RBLK, L, g);
          g.x *= ndy;
          g.y *= ndy;
          double2 g2 = make_double2(0.0, 0.0);
          const int arow = oF + (cc - c0) * nyb * RBLK;
#pragma unroll
          for (int y = 0; y < RNYB; ++y)
            if (y < nyb) mma2((y & 1) ? g2 : g, af[y], ldn(arow + y * RBLK, L));
          g.x += g2.x;
          g.y += g2.y;
          double cs[2][3];
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int tc = cc * 8 + 2 * L.q + e;
            const double Gv = e == 0 ? g.x : g.y;
            const bool cv = rv && IDX[tc] >= 0;
            if (cv && tc == tr) d0 += 0.5 * Gv;
            const bool off = cv && tc < tr;
            double k = e == 0 ? kv[j].x : kv[j].y;
            double gp[MAX_DX], gq[MAX_DX], gl[MAX_NLS];
            cov_grad<DFN, WFN, true>(xr, XS + tc * XD, cp, k, gp, gq, gl);
            const double Gm = off ? Gv : 0.0;
            s1 += off ? Gm * k : 0.0;
#pragma unroll
            for (int d = 0; d < 3; ++d) {
              rs[d] += off ? Gm * gp[d] : 0.0;
              cs[e][d] = off ? Gm * gq[d] : 0.0;
              thl[d] += off ? Gm * gl[d] : 0.0;
            }
          }
#pragma unroll
          for (int e = 0; e < 2; ++e)
#pragma unroll
            for (int d = 0; d < 3; ++d) {
              double v = cs[e][d];
              v += __shfl_xor_sync(0xffffffffu, v, 4);
              v += __shfl_xor_sync(0xffffffffu, v, 8);
              v += __shfl_xor_sync(0xffffffffu, v, 16);
              if (L.g == 0) c.colp[(long long)(rtri(r) + cc) * COLP + (2 * L.q + e) * 4 + d] = v;
            }
        }
        // row sums (8 x 3, at [g][d]) and the lengthscale partials (at [g][4 + d], summed over the quad)
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          double v = rs[d], u = thl[d];
          v += __shfl_xor_sync(0xffffffffu, v, 1);
          v += __shfl_xor_sync(0xffffffffu, v, 2);
          u += __shfl_xor_sync(0xffffffffu, u, 1);
          u += __shfl_xor_sync(0xffffffffu, u, 2);
          if (L.q == 0) {
            tp[L.g * 8 + d] = v;
            tp[L.g * 8 + 4 + d] = u;
          }
        }
      }
      // scalars of the task: fixed shuffle tree
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        d0 += __shfl_xor_sync(0xffffffffu, d0, o);
      }
      if (L.lane == 0) {
        tp[64] = s1;
        tp[65] = d0;
      }
    }
    rtrace(P, c, 22);
  }
  __syncthreads();
}

// ---- finalize: per-point sums over the task / block records in a fixed order; theta -------------------------
template <int DFN, int WFN>
static __device__ __noinline__ void ph_finalize(const ResParams& P, const Ctx& c) {
  const Lane L = make_lane();
  const int tid = threadIdx.x, w = L.w;
  const int nr = c.nr;
  const CovParams& cp = P.cp;
  const int* rowcnt = S_ROWCNT;
  const int FE = c.oXS;
  double th[MAX_NCOV];
#pragma unroll
  for (int t = 0; t < MAX_NCOV; ++t) th[t] = 0.0;
  const double coef = (WFN == WFN_SE) ? -2.0 : -3.0;         // w'(r)/r = coef * (k or k / (1 + sqrt3 r))
  // Two lanes per point: the even one sums the point's row records (8 doubles per task of its block row),
  // the odd one its column records (4 doubles per block below it) - all of them L2 round trips, issued
  // four records at a time; the pair is joined by a shuffle.  Fixed order => bit-reproducible.
  for (int t2 = tid; t2 < ((nr * 16 + 31) & ~31); t2 += RNT) {
    const int t = t2 >> 1, half = t2 & 1;
    const bool live = t < nr * 8;
    const int rb = live ? t >> 3 : 0, g0 = t & 7;
    double o[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (live && half == 0) {
      const int cnt = rowcnt[rb];
      const double* tp = c.taskp + ((long long)rb * MAXG) * TASKP + g0 * 8;
      int gi = 0;
      for (; gi + 4 <= cnt; gi += 4) {
        double2 v[4][4];
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
          for (int n = 0; n < 4; ++n) v[u][n] = *reinterpret_cast<const double2*>(tp + (long long)(gi + u) * TASKP + 2 * n);
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
          for (int n = 0; n < 4; ++n) {
            o[2 * n] += v[u][n].x;
            o[2 * n + 1] += v[u][n].y;
          }
      }
      for (; gi < cnt; ++gi)
#pragma unroll
        for (int n = 0; n < 8; ++n) o[n] += tp[(long long)gi * TASKP + n];
    } else if (live) {
      int r2 = rb;
      for (; r2 + 4 <= nr; r2 += 4) {
        double2 v[4][2];
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
          for (int n = 0; n < 2; ++n)
            v[u][n] = *reinterpret_cast<const double2*>(c.colp + (long long)(rtri(r2 + u) + rb) * COLP + g0 * 4 + 2 * n);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          o[0] += v[u][0].x;
          o[1] += v[u][0].y;
          o[2] += v[u][1].x;
          o[3] += v[u][1].y;
        }
      }
      for (; r2 < nr; ++r2) {
        const double* pc = c.colp + (long long)(rtri(r2) + rb) * COLP + g0 * 4;
#pragma unroll
        for (int n = 0; n < 4; ++n) o[n] += pc[n];
      }
    }
    double cc[4];
#pragma unroll
    for (int n = 0; n < 4; ++n) cc[n] = __shfl_xor_sync(0xffffffffu, o[n], 1);     // even lanes receive the column sums
    if (!live || half != 0) continue;
    double gxv[3];
    if (DFN == DFN_EUCLIDEAN) {
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        const double x = g_smem[FE + rb * RBLK + sw_off(g0, d)];
        gxv[d] = coef * cp.il2[d] * ((x * o[3] - o[d]) + (x * cc[3] - cc[d]));
        th[2 + d] += -coef * cp.il3[d] * (x * x * o[3] - 2.0 * x * o[d] + o[4 + d]);
      }
    } else {
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        gxv[d] = o[d] + cc[d];
        th[2 + d] += o[4 + d];
      }
    }
    c.gx[t * 3] = gxv[0];
    c.gx[t * 3 + 1] = gxv[1];
    c.gx[t * 3 + 2] = gxv[2];
  }
  // scalars of the tasks: lanes stride over the records in a fixed order
  for (int e = tid; e < nr * MAXG; e += RNT) {
    const int rb = e / MAXG, gi = e - rb * MAXG;
    if (gi < rowcnt[rb]) {
      th[1] += c.taskp[(long long)e * TASKP + 64];
      th[0] += c.taskp[(long long)e * TASKP + 65];
    }
  }
  // block reduction in a fixed order: shuffle tree, then the warps' partials in warp order
  double* red = g_smem + OFF_TS;                   // RNW x MAX_NCOV
#pragma unroll
  for (int t = 0; t < MAX_NCOV; ++t) {
    double v = th[t];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (L.lane == 0) red[w * MAX_NCOV + t] = v;
  }
  __syncthreads();
  if (tid < MAX_NCOV) {
    // theta = [nv, s2, l...]:  d/dnv = sum_p G_pp / 2;  d/ds2 = sum_p G_pp / 2 + sum_{q<p} G_pq k_pq / s2
    double v = 0.0, dg = 0.0;
    for (int i = 0; i < RNW; ++i) {
      v += red[i * MAX_NCOV + tid];
      dg += red[i * MAX_NCOV + 0];
    }
    if (tid == 1) v = v / cp.s2 + dg;
    P.gth_u[(long long)c.uid * MAX_NCOV + tid] = v;
  }
  __syncthreads();
}

// ---- one unit ---------------------------------------------------------------------------------------------------
template <int DFN, int WFN, class R1>
__device__ __forceinline__ void run_unit(const ResParams& P, const Ctx& c, Stage& stage) {
  const int tid = threadIdx.x;
  rtrace(P, c, 1);
  // W_i on its way into [R2 | F] (both still unused; res_class guarantees that all of it fits)
  if (c.pair && tid == 0 && !c.defer) tma_issue(stage, g_smem + c.oR2, c.pexp + EXP_W, rtri(c.ab));
  ph_gather<DFN>(P, c);
  rtrace(P, c, 33);
  if (c.pair) ph_lji<DFN, WFN, R1>(P, c, stage);
  dbg_dump(P, c, 1);
  rtrace(P, c, 2);
  ph_schur<DFN, WFN, R1>(P, c);
  dbg_dump(P, c, 2);
  rtrace(P, c, 3);
  // Z_i (all of it, or its first piece) travels into F while S is factored
  const int nyc = max(1, min(c.nyb, c.nF / max(c.ab, c.bb)));
  if (c.pair && tid == 0) {
    if (P.ready0) {              // started on the first flag: Z_i, alpha_i and the scalars come with the second
      const long long t0 = clock64();
      while (atomicAdd(P.ready + c.bi, 0) != P.epoch) {
        __nanosleep(100);
        if (clock64() - t0 > P.spin_limit) {
          atomicOr(P.status, ST_TIMEOUT);
          break;
        }
      }
      __threadfence();
    }
    tma_issue(stage, g_smem + c.oF, c.pexp + EXP_ZY, min(nyc, c.nyb) * c.ab);
  }
  ph_chol_inv(P, c);
  dbg_dump(P, c, 4);
  rtrace(P, c, 5);
  ph_ypart<R1>(P, c, stage, nyc);
  rtrace(P, c, 6);
  if (c.is_export) {
    // Everything this block's pairs need before THEIR gradient phase is exported: release them now
    // (measured: the CTAs that start with a pair waited 87 us for the whole block unit, 75 us of it
    // the parent's own work).  K^-1 follows in ph_grad under the second flag.
    __threadfence();
    __syncthreads();
    if (tid == 0) atomicExch(P.ready + c.bj, P.epoch);
  }
  if (!c.want_grad) {
    __syncthreads();
    return;
  }
  asm volatile("fence.proxy.async;\n" ::: "memory");     // Z_j / alpha_j / saved K: written with ordinary stores
  if (c.pair) {
    __syncthreads();
    // W_i (first piece) back into F for T
    if (tid == 0) {
      int r1 = 0, nb = 0;
      while (r1 < c.ab && nb + r1 + 1 <= c.nF) {
        nb += r1 + 1;
        ++r1;
      }
      tma_issue(stage, g_smem + c.oF, c.pexp + EXP_W, nb);
    }
    ph_t<R1>(P, c, stage);
    // own Z_j (all of it, or its first piece) into F for alpha_i, behind the V product
    if (tid == 0) tma_issue(stage, g_smem + c.oF, c.Zy, max(1, min(c.nyb, c.nF / c.bb)) * c.bb);
    dbg_dump(P, c, 6);
    rtrace(P, c, 7);
    ph_v<R1>(P, c);
    dbg_dump(P, c, 7);
    rtrace(P, c, 8);
    ph_alpha_i<R1>(P, c, stage);
    asm volatile("fence.proxy.async;\n" ::: "memory");   // alpha_i rows
  }
  if (c.pair) {                  // K_ii^-1 of the parent (its gradient phase): long done by now
    if (tid == 0) {
      const long long t0 = clock64();
      while (atomicAdd(P.ready2 + c.bi, 0) != P.epoch) {
        __nanosleep(100);
        if (clock64() - t0 > P.spin_limit) {
          atomicOr(P.status, ST_TIMEOUT);
          break;
        }
      }
      __threadfence();
    }
  }
  __syncthreads();
  rtrace(P, c, 9);
  ph_grad<DFN, WFN, R1>(P, c, stage);
  rtrace(P, c, 10);
  ph_finalize<DFN, WFN>(P, c);
  rtrace(P, c, 11);
  // The unit's scratch (saved covariance values, Z_j, alpha rows, per-task partials: ~340 KB per pair) is
  // dead now; without a hint every line of it is eventually written back from the L2 (measured: 124 MB
  // of DRAM writes per evaluation against 33 MB of algorithmic traffic).  discard.L2 drops the lines
  // instead.  (ph_finalize ended with a CTA barrier: every read is done; the next unit writes each
  // element before it reads it.)
  auto drop = [&](const double* base, int nlines) {
    for (int e = threadIdx.x; e < nlines; e += RNT)
      asm volatile("discard.global.L2 [%0], 128;" ::"l"(base + (size_t)e * 16) : "memory");
  };
  if (!c.is_export) {            // (a block unit keeps these in its exports)
    drop(c.Kji, c.bb * c.ab * 4);
    drop(c.Kjj, rtri(c.bb) * 4);
    drop(c.Zy, c.nyb * c.bb * 4);
    drop(c.Arow, c.nr * c.nyb * 4);
  }
  drop(c.colp, rtri(c.nr) * COLP / 16);
  drop(c.taskp, c.nr * MAXG * TASKP / 16);
}

// grid: persistent CTAs (<= one per SM, all co-resident: pairs spin on their parent's flag), RNT threads,
// R_SMEM_BYTES dynamic shared memory.
template <int DFN, int WFN>
__global__ void __launch_bounds__(RNT, 1) k_resident(ResParams Pk) {
  double* MISC = g_smem + OFF_MISC;
  Stage stage;
  stage.bar = reinterpret_cast<uint64_t*>(MISC);
  stage.par = 0;
  stage.status = Pk.status;
  stage.spin_limit = Pk.spin_limit;
  int* s_unit = reinterpret_cast<int*>(MISC + 4);
  Ctx* ctx = reinterpret_cast<Ctx*>(MISC + MISC_CTX);
  ResParams* Ps = reinterpret_cast<ResParams*>(MISC + MISC_PARAMS);
  if (threadIdx.x == 0) {
    mbar_init(stage.bar, 1);
    fence_mbar_init();
    *Ps = Pk;
    *S_TCUR = 0;
  }
  __syncthreads();
  const ResParams& P = *Ps;
  double* scratch = P.scratch + (long long)blockIdx.x * SCR_STRIDE;
  // Static unit lists (k_res_plan): block units first in every list, so a pair's parent never waits.
  const int slot_end = P.list_ptr[blockIdx.x + 1];
  for (int slot = P.list_ptr[blockIdx.x]; slot < slot_end; ++slot) {
    const int uid = P.order[slot];
    int bi = -1, bj = uid;
    if (uid >= P.B) {
      bi = P.edges[2 * (uid - P.B)];
      bj = P.edges[2 * (uid - P.B) + 1];
    }
    const long long ja = P.block_ptr[bj];
    const int b = (int)(P.block_ptr[bj + 1] - ja);
    long long ia = 0;
    int a = 0;
    if (bi >= 0) {
      ia = P.block_ptr[bi];
      a = (int)(P.block_ptr[bi + 1] - ia);
    }
    int defer = 0;                  // (thread 0)
    if (bi >= 0) {
      // the parent block's unit (earlier in the queue, possibly still running on another SM)
      if (threadIdx.x == 0) {
        // (a pair with an empty second block copies block i's final results: second flag)
        const int* flag = (b == 0 ? P.ready2 : (P.ready0 ? P.ready0 : P.ready)) + bi;
        if (b != 0 && P.ready0 && P.defer_ok && atomicAdd(const_cast<int*>(flag), 0) != P.epoch) {
          // Not released yet: the pair starts anyway - its coordinate records and K_ji do not need the
          // parent - and waits in ph_lji, just before it requests W_i.
          defer = 1;
        } else {
          const long long t0 = clock64();
          while (atomicAdd(const_cast<int*>(flag), 0) != P.epoch) {
            __nanosleep(100);
            if (clock64() - t0 > P.spin_limit) {
              atomicOr(P.status, ST_TIMEOUT);
              break;
            }
          }
          __threadfence();
        }
      }
      __syncthreads();
    }
    if (bi >= 0 && b == 0) {
      // pair with an empty second block: the unit IS block i
      const long long src = bi;
      if (threadIdx.x == 0) P.ll_u[uid] = P.ll_u[src];
      if (threadIdx.x < MAX_NCOV) P.gth_u[(long long)uid * MAX_NCOV + threadIdx.x] = P.gth_u[src * MAX_NCOV + threadIdx.x];
      for (int e = threadIdx.x; e < GX_STRIDE; e += RNT) P.gx_u[(long long)uid * GX_STRIDE + e] = P.gx_u[src * GX_STRIDE + e];
      continue;
    }
    const int ab = (a + 7) >> 3, bb = (b + 7) >> 3;
    const int cls = res_class(ab, bb);
    if (cls == 2) {
      if (threadIdx.x == 0) {
        atomicOr(P.status, ST_OVERFLOW);
        if (bi < 0) {                                        // nobody may wait forever; the tile pipeline redoes it
          if (P.ready0) atomicExch(P.ready0 + bj, P.epoch);
          atomicExch(P.ready + bj, P.epoch);
          atomicExch(P.ready2 + bj, P.epoch);
        }
      }
      continue;
    }
    if (b == 0) {                  // empty block unit
      if (threadIdx.x == 0) {
        P.ll_u[uid] = 0.0;
        double* oexp = P.exports + (long long)bj * EXP_STRIDE;
        oexp[EXP_SCAL + 0] = 0.0;
        oexp[EXP_SCAL + 1] = 0.0;
      }
      if (threadIdx.x < MAX_NCOV) P.gth_u[(long long)uid * MAX_NCOV + threadIdx.x] = 0.0;
      __threadfence();
      __syncthreads();
      if (threadIdx.x == 0) {
        if (P.ready0) atomicExch(P.ready0 + bj, P.epoch);
        atomicExch(P.ready + bj, P.epoch);
        atomicExch(P.ready2 + bj, P.epoch);
      }
      continue;
    }
    const bool r1g = cls == 1;
    if (threadIdx.x == 0) {
      Ctx& c = *ctx;
      c.uid = uid; c.bi = bi; c.bj = bj; c.a = a; c.b = b; c.ab = ab; c.bb = bb; c.nr = ab + bb;
      c.nyb = P.nyb; c.pair = ab > 0; c.is_export = bi < 0; c.want_grad = P.want_grad; c.dx = P.dx; c.dy = P.dy;
      c.r1g = r1g;
      c.defer = defer;
      c.ia = ia; c.ja = ja;
      c.oXS = OFF_XS;
      c.oR1 = OFF_XS + res_xs_blocks(ab, bb) * RBLK;
      c.oR2 = c.oR1 + (r1g ? 0 : bb * ab * RBLK);
      c.oF = c.oR2 + rtri(bb) * RBLK;
      c.nF = res_free_blocks(ab, bb, r1g);
      c.R1g = scratch + SCR_R1;
      c.pexp = c.pair ? P.exports + (long long)bi * EXP_STRIDE : nullptr;
      c.oexp = c.is_export ? P.exports + (long long)bj * EXP_STRIDE : nullptr;
      c.Zy = c.is_export ? c.oexp + EXP_ZY : scratch + SCR_ZY;
      c.Arow = c.is_export ? c.oexp + EXP_AROW : scratch + SCR_AROW;
      c.Kjj = c.is_export ? c.oexp + EXP_KSAVE : scratch + SCR_KJJ;
      c.Kji = scratch + SCR_KJI;
      c.colp = scratch + SCR_COLP;
      c.taskp = scratch + SCR_TASKP;
      c.gx = P.gx_u + (long long)uid * GX_STRIDE;
    }
    __syncthreads();
    if (r1g) run_unit<DFN, WFN, R1Glob>(P, *ctx, stage);
    else run_unit<DFN, WFN, R1Smem>(P, *ctx, stage);
    if (bi < 0) {                  // all exports complete
      __threadfence();
      __syncthreads();
      if (threadIdx.x == 0) atomicExch(P.ready2 + bj, P.epoch);
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// Launch plan on the device (one CTA): which blocks have to be factored, the pair units in
// descending size order (longest first on the dynamic queue), the fit check.  No host round trip:
// block sizes never leave the GPU.
constexpr int PLAN_NKEY = 3 * BMAXB + 5 * BMAXB + 1;       // values of the pairs' cost key 3 ab + 5 bb
// shared memory of res_plan_body for a CTA of nt threads: need[B] | key[E] | histograms | bucket starts
__host__ __device__ inline size_t res_plan_smem(int B, int E, int nt) {
  return ((size_t)B + E + (size_t)(nt / 32 + 1) * PLAN_NKEY + 1) * sizeof(int);
}
struct PlanParams {
  const long long* block_ptr;
  const int* edges;
  const unsigned char* active;     // per unit (B + E), or nullptr = all
  int B, E;
  int G;                           // CTAs of the resident launch
  int sort_blocks;                 // block units in size order (GPRF_RES_SORTBLK=0: in id order)
  int* order;                      // out: the CTAs' unit lists, back to back
  int* list_ptr;                   // out: G + 1 list bounds
  int* counts;                     // out: [n units, n block units, -, -]
  int* status;                     // out (reset here)
  int* info;                       // per unit, reset here
};

#ifndef GPRF_RES_KERNEL_ONLY
// One CTA (a multiple of 32 threads); sh: res_plan_smem(B, E, blockDim.x) bytes of shared memory.  Also called at the end of the single-CTA
// bucketing kernel of partition.cuh (k_bucket_small), which saves a launch per evaluation.
__device__ __forceinline__ void res_plan_body(const PlanParams& Q, int* sh) {
  int* need = sh;
  int* key = sh + Q.B;
  const int tid = threadIdx.x, nt = blockDim.x;
  __shared__ int s_over, s_nb, s_np;
  __shared__ int s_bstart[BMAXB + 2];
  if (tid == 0) {
    s_over = 0;
    s_nb = 0;
    s_np = 0;
  }
  for (int bq = tid; bq < Q.B; bq += nt) need[bq] = (!Q.active || Q.active[bq]) ? 1 : 0;
  for (int u = tid; u < Q.B + Q.E; u += nt) Q.info[u] = 0;
  __syncthreads();
  for (int e = tid; e < Q.E; e += nt) {
    const bool act = !Q.active || Q.active[Q.B + e];
    const int i = Q.edges[2 * e], j = Q.edges[2 * e + 1];
    const int a = (int)(Q.block_ptr[i + 1] - Q.block_ptr[i]);
    const int b = (int)(Q.block_ptr[j + 1] - Q.block_ptr[j]);
    // cost key: measured pair times fit 11.5 ab + 19.0 bb - 228 us (8-blocks of block i / block j)
    // (clamped: a structure with oversized blocks is handed to the tile pipeline, but the plan still runs)
    key[e] = act ? (b > 0 ? min(PLAN_NKEY - 1, 3 * ((a + 7) >> 3) + 5 * ((b + 7) >> 3)) : 0) : -1;
    if (act) {
      need[i] = 1;                      // benign race: everybody writes 1
      if (b > 0 && res_class((a + 7) >> 3, (b + 7) >> 3) == 2) s_over = 1;
    }
  }
  __syncthreads();
  for (int bq = tid; bq < Q.B; bq += nt) {
    const int s = (int)(Q.block_ptr[bq + 1] - Q.block_ptr[bq]);
    if (need[bq] && res_class(0, (s + 7) >> 3) == 2) s_over = 1;
  }
  __syncthreads();
  // Static unit lists, one per CTA of the resident launch.  Measured on the README configuration the
  // dynamic largest-first queue left 46 CTAs with three ~165 us pair units and 102 with two (the
  // blocks take ~75 us on 100 CTAs while the other 48 wait for them, so every CTA is "busy" until
  // then): 545 us against a mean load of 430.  Here the CTAs that must take one pair more get the
  // SMALLEST pairs, and inside each class the pairs are dealt in snake order (large with small).
  //   blocks: the k-th needed block in (size class, id) order -> CTA k % G, slot k / G
  //   pairs by rank r (size descending): q = np / G, rem = np % G; the first (G - rem) q ranks go to
  //   CTAs 0 .. G-rem-1 (q each), the rest to CTAs G-rem .. G-1 (q + 1 each)
  if (tid < 32) {
    // need[bq] = 1 + the block's index among the needed blocks in (size class ascending, id ascending) order: the
    // CTAs are filled from 0 upwards and CTA 0 also gets the LARGEST pairs (below), so the smallest blocks
    // go where the pairs are heaviest (the block-carrying CTAs with the top pairs end the launch: measured
    // 493 us against 451 for the mean block-carrying CTA; a block unit takes 46 + 2.5 us per 8 points).
    constexpr int NCLS = BMAXB + 2;                  // size classes 0 .. BMAXB, oversized
    for (int v = tid; v < NCLS; v += 32) s_bstart[v] = 0;
    __syncwarp();
    auto cls = [&](int bq) {
      return Q.sort_blocks ? min(NCLS - 1, (int)((Q.block_ptr[bq + 1] - Q.block_ptr[bq] + 7) >> 3)) : 0;
    };
    for (int b0 = 0; b0 < Q.B; b0 += 32) {
      const int bq = b0 + tid;
      if (bq < Q.B && need[bq]) atomicAdd(&s_bstart[cls(bq)], 1);
    }
    __syncwarp();
    if (tid == 0) {
      int at = 0;
      for (int v = 0; v < NCLS; ++v) {
        const int cnt = s_bstart[v];
        s_bstart[v] = at;
        at += cnt;
      }
      s_nb = at;
    }
    __syncwarp();
    for (int b0 = 0; b0 < Q.B; b0 += 32) {
      const int bq = b0 + tid;
      const bool f = bq < Q.B && need[bq];
      const int v = f ? cls(bq) : 0;
      unsigned grp = __ballot_sync(0xffffffffu, f);              // lanes with a needed block of the same class
      for (int bit = 0; bit < 5; ++bit) {
        const bool one = (v >> bit) & 1;
        const unsigned mb = __ballot_sync(0xffffffffu, one);
        grp &= one ? mb : ~mb;
      }
      if (!f) grp = 1u << tid;
      const int leader = __ffs(grp) - 1;
      int base = 0;
      if (f && tid == leader) {
        base = s_bstart[v];
        s_bstart[v] = base + __popc(grp);
      }
      base = __shfl_sync(0xffffffffu, base, leader);
      if (f) need[bq] = 1 + base + __popc(grp & ((1u << tid) - 1u));
      __syncwarp();
    }
  }
  {
    int mine = 0;
    for (int e = tid; e < Q.E; e += nt) mine += key[e] >= 0 ? 1 : 0;
    for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
    if ((tid & 31) == 0 && mine) atomicAdd(&s_np, mine);
  }
  __syncthreads();
  const int G = Q.G, nb = s_nb, np_ = s_np;
  const int kb = nb / G, rb = nb % G, q = np_ / G, rem = np_ % G, mB = G - rem;
  auto nblk = [&](int w) { return kb + (w < rb ? 1 : 0); };
  auto lptr = [&](int w) { return w * kb + min(w, rb) + w * q + max(0, w - mB); };
  for (int w = tid; w <= G; w += nt) Q.list_ptr[w] = lptr(w);
  for (int bq = tid; bq < Q.B; bq += nt) {
    if (need[bq]) {
      const int k = need[bq] - 1, w = k % G;
      Q.order[lptr(w) + k / G] = bq;
    }
  }
  // pairs by (cost key descending, id ascending): the key takes at most PLAN_NKEY values, so a stable
  // counting sort does it in O(E) (the O(E^2) rank sort it replaces was 45 000 of the kernel's 60 000
  // warp instructions at E = 342, ~10 us on the one SM that runs the plan).  Warp v owns the contiguous
  // edge range [v Le, (v + 1) Le): per-warp histograms, per bucket a prefix over the warps, the bucket
  // starts, then every warp places its edges in order (ranks inside a group of 32 by __match_any_sync).
  {
    int* kh = sh + Q.B + Q.E;                       // [nwarp][PLAN_NKEY] counts -> prefixes
    const int nwarp = nt >> 5, wv = tid >> 5, lane = tid & 31;
    int* kstart = kh + nwarp * PLAN_NKEY;           // [PLAN_NKEY + 1]
    for (int e = tid; e < nwarp * PLAN_NKEY; e += nt) kh[e] = 0;
    __syncthreads();
    const int Le = (Q.E + nwarp - 1) / nwarp;
    const int e0 = wv * Le, e1 = min(Q.E, e0 + Le);
    for (int e = e0 + lane; e < e1; e += 32)
      if (key[e] >= 0) atomicAdd(&kh[wv * PLAN_NKEY + (PLAN_NKEY - 1 - key[e])], 1);     // bucket 0 = largest key
    __syncthreads();
    for (int kb2 = tid; kb2 < PLAN_NKEY; kb2 += nt) {
      int cnt = 0;
      for (int v = 0; v < nwarp; ++v) {
        const int x = kh[v * PLAN_NKEY + kb2];
        kh[v * PLAN_NKEY + kb2] = cnt;
        cnt += x;
      }
      kstart[kb2] = cnt;
    }
    __syncthreads();
    if (tid == 0) {                                  // PLAN_NKEY = 161 totals -> starts
      int at = 0;
      for (int kb2 = 0; kb2 < PLAN_NKEY; ++kb2) {
        const int cnt = kstart[kb2];
        kstart[kb2] = at;
        at += cnt;
      }
    }
    __syncthreads();
    for (int q0 = e0; q0 < e1; q0 += 32) {
      const int e = q0 + lane;
      const int ke = e < e1 ? key[e] : -1;
      const bool live = ke >= 0;
      const int bk = live ? PLAN_NKEY - 1 - ke : -1 - lane;
      const unsigned grp = __match_any_sync(0xffffffffu, bk);
      const int within = __popc(grp & ((1u << lane) - 1u));
      const int leader = __ffs(grp) - 1;
      int base = 0;
      if (live && lane == leader) {
        base = kh[wv * PLAN_NKEY + bk];
        kh[wv * PLAN_NKEY + bk] = base + __popc(grp);
      }
      base = __shfl_sync(0xffffffffu, base, leader);
      if (live) {
        const int rank = kstart[bk] + base + within;
        int w, round;
        if (rank < mB * q) {
          round = rank / mB;
          const int pos = rank - round * mB;
          w = (round & 1) ? mB - 1 - pos : pos;
        } else {
          const int r2 = rank - mB * q;
          round = r2 / rem;
          const int pos = r2 - round * rem;
          w = mB + ((round & 1) ? rem - 1 - pos : pos);
        }
        Q.order[lptr(w) + nblk(w) + round] = Q.B + e;
      }
      __syncwarp();
    }
  }
  // A CTA without a block unit starts with a pair and waits for that pair's parent: put the pair with the
  // smallest parent block (the first to be released) in front.  These CTAs carry one pair more than the
  // others and finish last, so what they wait at the start is on the launch's critical path.
  __threadfence_block();
  __syncthreads();
  for (int w = tid; w < G; w += nt) {
    if (nblk(w) != 0) continue;
    const int beg = lptr(w), cnt = lptr(w + 1) - beg;
    int best = 0, bs = 0x7fffffff;
    for (int k = 0; k < cnt; ++k) {
      const int e = Q.order[beg + k] - Q.B;
      const int i = Q.edges[2 * e];
      const int sz = (int)(Q.block_ptr[i + 1] - Q.block_ptr[i]);
      if (sz < bs) {
        bs = sz;
        best = k;
      }
    }
    if (best != 0) {
      const int t0 = Q.order[beg];
      Q.order[beg] = Q.order[beg + best];
      Q.order[beg + best] = t0;
    }
  }
  __syncthreads();
  if (tid == 0) {
    Q.counts[0] = s_nb + s_np;
    Q.counts[1] = s_nb;
    Q.counts[2] = 0;
    Q.counts[3] = 0;
    *Q.status = s_over ? ST_OVERFLOW : 0;
  }
}

__global__ void k_res_plan(PlanParams Q) {
  extern __shared__ int sh[];          // res_plan_smem(B, E, blockDim.x) bytes
  res_plan_body(Q, sh);
}

#endif  // GPRF_RES_KERNEL_ONLY

// ---------------------------------------------------------------------------------------------------
// combine (gprf.py:245-291) over the resident path's per-unit results
struct ResCombine {
  const long long* perm;
  const int* pos_block;
  const long long* block_ptr;
  const int* adj_ptr;
  const int* adj_edge;
  const int* adj_side;
  const int* edges;
  const int* deg;                  // per block
  int raw;                         // masked evaluation with unit weights (gprf_set_unit_mask)
  const unsigned char* active;     // per unit or nullptr
  const double* ll_u;
  const double* gth_u;
  const double* gx_u;
  const int* status;
  int B, E, dx;
  long long plen;
};

// blocks [1, ...): one thread per perm position; block 0: the scalars (ll, grad theta) and the status
#ifndef GPRF_RES_KERNEL_ONLY
__global__ void k_res_combine(ResCombine C, double* out, int want_gx, int want_cov, double* status_out) {
  if (blockIdx.x == 0) {
    __shared__ double red[256][1 + MAX_NCOV];
    const int tid = threadIdx.x;
    double v[1 + MAX_NCOV];
#pragma unroll
    for (int t = 0; t < 1 + MAX_NCOV; ++t) v[t] = 0.0;
    for (int u = tid; u < C.B + C.E; u += 256) {
      if (C.active && !C.active[u]) continue;
      long long s;
      double wgt;
      if (u < C.B) {
        s = C.block_ptr[u + 1] - C.block_ptr[u];
        wgt = C.raw ? 1.0 : 1.0 - (double)C.deg[u];
      } else {
        const int i = C.edges[2 * (u - C.B)], j = C.edges[2 * (u - C.B) + 1];
        s = (C.block_ptr[i + 1] - C.block_ptr[i]) + (C.block_ptr[j + 1] - C.block_ptr[j]);
        wgt = 1.0;
      }
      if (s == 0) continue;
      v[0] += wgt * C.ll_u[u];
      if (want_cov)
#pragma unroll
        for (int t = 0; t < MAX_NCOV; ++t) v[1 + t] += wgt * C.gth_u[(long long)u * MAX_NCOV + t];
    }
#pragma unroll
    for (int t = 0; t < 1 + MAX_NCOV; ++t) red[tid][t] = v[t];
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
      if (tid < o)
#pragma unroll
        for (int t = 0; t < 1 + MAX_NCOV; ++t) red[tid][t] += red[tid + o][t];
      __syncthreads();
    }
    if (tid < 1 + MAX_NCOV) out[tid] = red[0][tid];
    if (tid == 0 && status_out) *status_out = (double)*C.status;
    return;
  }
  if (!want_gx) return;
  const long long pos = (long long)(blockIdx.x - 1) * blockDim.x + threadIdx.x;
  if (pos >= C.plen) return;
  const int bq = C.pos_block[pos];
  const int lp = (int)(pos - C.block_ptr[bq]);
  double g[3] = {0.0, 0.0, 0.0};
  if (!C.active || C.active[bq]) {
    const double wgt = C.raw ? 1.0 : 1.0 - (double)C.deg[bq];
    const double* p = C.gx_u + (long long)bq * GX_STRIDE + (long long)lp * 3;
    g[0] += wgt * p[0];
    g[1] += wgt * p[1];
    g[2] += wgt * p[2];
  }
  // The block's edges, four at a time: every stage of the chain  adjacency -> edge -> size of block i -> row of
  // the pair's gradient  is loaded for all four before the next stage starts; the sums keep the edge order.
  // (Measured: the kernel stays at 12.7 us - CTA 0's walk over the unit records is what bounds it.)
  const int a_end = C.adj_ptr[bq + 1];
  for (int a0 = C.adj_ptr[bq]; a0 < a_end; a0 += 4) {
    bool on[4];
    int e[4], ib[4], off[4];
    double v[4][3];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      on[u] = a0 + u < a_end;
      e[u] = on[u] ? C.adj_edge[a0 + u] : 0;
      ib[u] = (on[u] && C.adj_side[a0 + u]) ? 0 : -1;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (on[u] && C.active && !C.active[C.B + e[u]]) on[u] = false;
      if (on[u] && ib[u] == 0) ib[u] = C.edges[2 * e[u]];
      else ib[u] = -1;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
      off[u] = ib[u] >= 0 ? (((int)(C.block_ptr[ib[u] + 1] - C.block_ptr[ib[u]]) + 7) >> 3) * 8 : 0;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const double* p = C.gx_u + (long long)(C.B + e[u]) * GX_STRIDE + (long long)(off[u] + lp) * 3;
      v[u][0] = on[u] ? p[0] : 0.0;
      v[u][1] = on[u] ? p[1] : 0.0;
      v[u][2] = on[u] ? p[2] : 0.0;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (on[u]) {
        g[0] += v[u][0];
        g[1] += v[u][1];
        g[2] += v[u][2];
      }
    }
  }
  const long long n = C.perm[pos];
  for (int d = 0; d < C.dx; ++d) out[1 + MAX_NCOV + n * C.dx + d] = g[d];
}
#endif  // GPRF_RES_KERNEL_ONLY

}  // namespace res
}  // namespace gprf
