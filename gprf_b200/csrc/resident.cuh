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
constexpr int R_MISC_DOUBLES = 320;
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
constexpr long long EXP_AROW = EXP_ZY + (long long)RNYB * EMAXB * RBLK;     // k * RNYB + yb
constexpr long long EXP_SCAL = EXP_AROW + (long long)EMAXB * RNYB * RBLK;   // logdet, |Z|^2
constexpr long long EXP_STRIDE = EXP_SCAL + 16;
// per-CTA scratch (doubles)
constexpr long long SCR_ZY = 0;                                              // yb * bb + k
constexpr long long SCR_AROW = (long long)RNYB * EMAXB * RBLK;              // 2*EMAXB rows x RNYB
constexpr long long SCR_COLP = SCR_AROW + 2LL * EMAXB * RNYB * RBLK;
constexpr int COLP = 24;                                                     // per G block: column sums 8 x 3
constexpr long long SCR_TASKP = SCR_COLP + (long long)(2 * EMAXB) * (2 * EMAXB + 1) / 2 * COLP;
constexpr int TASKP = 32;                                                    // per task: row sums 8 x 3, theta 5
constexpr int GCOLS = 4;                                                     // G blocks per task
constexpr int MAXG = (2 * EMAXB + GCOLS - 1) / GCOLS;                        // column groups per row
constexpr long long SCR_KJI = SCR_TASKP + (long long)(2 * EMAXB) * MAXG * TASKP;   // saved K_ji: row * ab + c
constexpr long long SCR_KJJ = SCR_KJI + (long long)EMAXB * EMAXB * RBLK;           // saved K_jj: lower packed
constexpr long long SCR_R1 = SCR_KJJ + (long long)RTRI * RBLK;                      // R1 of units too big for smem
constexpr long long SCR_STRIDE = SCR_R1 + (long long)EMAXB * EMAXB * RBLK;
constexpr int GX_STRIDE = 2 * EMAXB * 8 * 3;           // per-unit gradX rows (padded local order)

enum { ST_OVERFLOW = 1, ST_NOTPD = 2 };

__host__ __device__ __forceinline__ int rtri(int i) { return i * (i + 1) / 2; }
// free staging blocks behind XS / R1 / R2 (negative: does not fit)
__host__ __device__ __forceinline__ int res_free_blocks(int ab, int bb, bool r1_global) {
  const int used = (ab + bb) * 8 * XD + ((r1_global ? 0 : bb * ab) + rtri(bb)) * RBLK;
  return (R_CAP_DOUBLES - used) / RBLK - (R_CAP_DOUBLES < used ? 1 : 0);
}
// 0: everything in shared memory; 1: R1 (the bb x ab coupling matrix) in this CTA's L2-resident
// scratch, the rest in shared memory; 2: does not fit (tile pipeline).
__host__ __device__ __forceinline__ int res_class(int ab, int bb) {
  if (ab > BMAXB || bb > BMAXB) return 2;
  if (res_free_blocks(ab, bb, false) >= R_MIN_STAGE_BLOCKS) return 0;
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
  int* counter;                 // dynamic queue head
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
  int dbg_unit, dbg_phase;      // debug dump of R1 / R2 after a phase (-1: off)
  double* dbg_out;              // 2 x (160 x 160) doubles
  unsigned long long* trace;    // debug timeline (RTRACE_SLOTS (tag, ns) pairs per CTA) or nullptr
};

// ---- swizzled 8x8 block ---------------------------------------------------------------------
// element (r, c) lives at  8 * (r ^ ((r >> 1) & 1)) + (c ^ (r & 4)).
__host__ __device__ __forceinline__ int sw_off(int r, int c) { return ((r ^ ((r >> 1) & 1)) << 3) + (c ^ (r & 4)); }

struct Lane {
  int w, lane, g, q;
  int on, ot0, ot1;             // offsets of the row-major / transposed fragment elements
};
__device__ __forceinline__ Lane make_lane() {
  Lane L;
  L.w = threadIdx.x >> 5;
  L.lane = threadIdx.x & 31;
  L.g = L.lane >> 2;
  L.q = L.lane & 3;
  L.on = sw_off(L.g, 2 * L.q);
  L.ot0 = sw_off(2 * L.q, L.g);
  L.ot1 = sw_off(2 * L.q + 1, L.g);
  // opaque to the optimiser: under register pressure it would otherwise recompute these from
  // %tid at every use (measured: 13 % of all executed instructions)
  asm volatile("" : "+r"(L.on), "+r"(L.ot0), "+r"(L.ot1), "+r"(L.g), "+r"(L.q));
  return L;
}
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
};
__device__ __forceinline__ void tma_issue(const Stage& S, double* dst, const double* src, int nblk) {
  fence_proxy_async();
  const uint32_t bytes = (uint32_t)nblk * RBLK * 8;
  mbar_expect_tx(S.bar, bytes);
  bulk_g2s(dst, src, bytes, S.bar);
}
__device__ __forceinline__ void tma_wait(Stage& S) {
  mbar_wait(S.bar, S.par & 1u);
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

// ---- per-unit geometry ---------------------------------------------------------------------------
struct Unit {
  int uid, bi, bj;              // bi < 0: block unit
  int a, b, ab, bb;             // points / 8-blocks of the i part (0 for block units) and the j part
  long long ia, ja;             // offsets of the two blocks in perm
};

// coordinates of local row t (i part: [0, 8 ab), j part: [8 ab, 8 ab + 8 bb))
__device__ __forceinline__ const double* xs_row(const double* XS, int t) { return XS + t * XD; }

constexpr int RTRACE_SLOTS = 512;
// Debug timeline: thread 0 appends (unit << 16 | tag, %globaltimer) to this CTA's slots.
__device__ __forceinline__ void rtrace(const ResParams& P, int* cursor, int uid, int tag) {
  if (P.trace && threadIdx.x == 0 && *cursor < RTRACE_SLOTS) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    unsigned long long* wp = P.trace + ((long long)blockIdx.x * RTRACE_SLOTS + *cursor) * 2;
    wp[0] = ((unsigned long long)uid << 16) | (unsigned long long)tag;
    wp[1] = t;
    ++*cursor;
  }
}

// ---------------------------------------------------------------------------------------------------
// MB: size of the per-warp register arrays = upper bound on ab and bb (RMAXB for ordinary units,
// BMAXB for the few big ones).  Row-owned phases deal rows to the 16 warps RNW at a time; units
// with more than 16 block rows take a second pass.
template <int DFN, int WFN, int MB>
__device__ __forceinline__ void run_unit(const ResParams& P, const Unit& u, double* smem, Stage& stage,
                                         double* scratch, bool r1_global) {
  const Lane L = make_lane();
  const int tid = threadIdx.x;
  const int w = L.w;
  const int a = u.a, b = u.b, ab = u.ab, bb = u.bb;
  const int nr = ab + bb;                          // block rows of the whole unit
  const int nyb = P.nyb;
  const bool pair = ab > 0;
  const bool is_export = u.bi < 0;                 // block units export for their pairs
  double* WB = smem;
  double* MISC = WB + R_WB_BLOCKS * RBLK;
  double* XS = MISC + R_MISC_DOUBLES;
  double* R1 = r1_global ? scratch + SCR_R1 : XS + nr * 8 * XD;
  double* R2 = XS + nr * 8 * XD + (r1_global ? 0 : bb * ab * RBLK);
  double* F = R2 + rtri(bb) * RBLK;                // staging area behind the matrices
  const int nF = res_free_blocks(ab, bb, r1_global);
  int* s_fail = reinterpret_cast<int*>(MISC + 4) + 1;
  int* s_tcur = reinterpret_cast<int*>(MISC + 5);
  int* s_task = reinterpret_cast<int*>(MISC + 5) + 1;
  int* s_ntask = reinterpret_cast<int*>(MISC + 6);
  double* s_q = MISC + 8;                          // [RNW]
  double* s_ld = MISC + 104;                       // [RNW]
  int* IDX = reinterpret_cast<int*>(MISC + 120);   // [2 * EMAXB * 8] global point index or -1
  double* WD = WB;                                 // diagonal-block inverses during the factorisation
  const double* pexp = pair ? P.exports + (long long)u.bi * EXP_STRIDE : nullptr;
  double* oexp = is_export ? P.exports + (long long)u.bj * EXP_STRIDE : nullptr;
  double* Zy = is_export ? oexp + EXP_ZY : scratch + SCR_ZY;
  double* Arow = is_export ? oexp + EXP_AROW : scratch + SCR_AROW;
  double* Kjj = is_export ? oexp + EXP_KSAVE : scratch + SCR_KJJ;
  double* Kji = scratch + SCR_KJI;
  double* colp = scratch + SCR_COLP;
  double* taskp = scratch + SCR_TASKP;
  double* gx = P.gx_u + (long long)u.uid * GX_STRIDE;
  const CovParams& cp = P.cp;
  rtrace(P, s_tcur, u.uid, 1);

  auto dbg_dump = [&](int phase) {
    if (P.dbg_unit != u.uid || P.dbg_phase != phase) return;
    __syncthreads();
    for (int e = tid; e < 160 * 160; e += RNT) {
      const int r = e / 160, c = e % 160;
      double v1 = 0.0, v2 = 0.0;
      if (pair && r < bb * 8 && c < ab * 8) v1 = R1[((r >> 3) * ab + (c >> 3)) * RBLK + sw_off(r & 7, c & 7)];
      if (r < bb * 8 && c < bb * 8 && (c >> 3) <= (r >> 3))
        v2 = R2[(rtri(r >> 3) + (c >> 3)) * RBLK + sw_off(r & 7, c & 7)];
      P.dbg_out[e] = v1;
      P.dbg_out[160 * 160 + e] = v2;
    }
    __syncthreads();
  };
  auto tri_len = [](int r) { return r + 1; };
  auto tri_pos = [](int r) { return rtri(r); };

  // ---- P0: W_i on its way into [R2 | F] (both still unused), gather coordinates ----------------------
  const int capW1 = rtri(bb) + nF;                 // staging capacity for W_i before S exists
  if (pair && tid == 0) tma_issue(stage, R2, pexp + EXP_W, first_piece(capW1, ab, tri_len));
  if (tid == 0) *s_fail = 0;
  for (int t = tid; t < nr * 8; t += RNT) {
    long long idx = -1;
    if (t < ab * 8) {
      if (t < a) idx = P.perm[u.ia + t];
    } else if (t - ab * 8 < b) {
      idx = P.perm[u.ja + (t - ab * 8)];
    }
    IDX[t] = (int)idx;
    double rec[XD];
#pragma unroll
    for (int d = 0; d < MAX_DX + 1; ++d) rec[d] = (idx >= 0 && d < P.dx) ? P.X[idx * P.dx + d] : 0.0;
    point_terms(DFN, rec);
#pragma unroll
    for (int d = 0; d < XD; ++d) XS[t * XD + d] = rec[d];
  }
  __syncthreads();

  // covariance fragment of block (rb, cb) in local block coordinates: rows 8 rb + g, cols 8 cb + 2q, +1
  auto cov_frag = [&](int rb, int cb, bool with_diag) {
    const int tr = rb * 8 + L.g;
    const int tc = cb * 8 + 2 * L.q;
    const bool rv = IDX[tr] >= 0;
    const double* xr = xs_row(XS, tr);
    double v0 = cov_value<DFN, WFN>(xr, xs_row(XS, tc), cp);
    double v1 = cov_value<DFN, WFN>(xr, xs_row(XS, tc + 1), cp);
    v0 = (rv && IDX[tc] >= 0) ? v0 : 0.0;
    v1 = (rv && IDX[tc + 1] >= 0) ? v1 : 0.0;
    if (with_diag) {
      if (tr == tc) v0 = rv ? cp.s2 + cp.nv : 1.0;
      if (tr == tc + 1) v1 = rv ? cp.s2 + cp.nv : 1.0;
    }
    return make_double2(v0, v1);
  };

  // ---- P1: L_ji = K_ji W_i^T  (one row of L_ji per warp, its K_ji fragments in registers) ----------------
  if (pair) {
    bool ahead = true;
    for (int rbase = 0; rbase < bb; rbase += RNW) {
      const int row = rbase + w;
      double2 kf[MB];
      if (row < bb) {
#pragma unroll
        for (int k = 0; k < MB; ++k) {
          kf[k] = (k < ab) ? cov_frag(ab + row, k, false) : make_double2(0.0, 0.0);
          if (k < ab && P.want_grad) stn(Kji + (row * ab + k) * RBLK, L, kf[k]);   // saved for the gradient
        }
      }
      staged_rows(stage, R2, capW1, pexp + EXP_W, ab, tri_len, tri_pos, ahead, [&](int c0, int c1, const double* base) {
        if (row < bb) {
          for (int c = c0; c < c1; ++c) {
            const double* wrow = base + (rtri(c) - rtri(c0)) * RBLK;
            double2 acc0 = make_double2(0.0, 0.0), acc1 = make_double2(0.0, 0.0);
            for_desc<MB>(c + 1, [&](auto kc) {
              constexpr int k = decltype(kc)::value;
              mma2((k & 1) ? acc1 : acc0, kf[k], ldn(wrow + k * RBLK, L));
            });
            stn(R1 + (row * ab + c) * RBLK, L, make_double2(acc0.x + acc1.x, acc0.y + acc1.y));
          }
        }
      });
      ahead = false;
    }
  }
  dbg_dump(1);
  rtrace(P, s_tcur, u.uid, 2);

  // ---- P2: S = K_jj + nv I - L_ji L_ji^T (lower blocks; a row's columns split by parity over two tasks) ----
  __syncthreads();
  for (int t = w; t < bb; t += RNW) {
#pragma unroll 1
    for (int part = 0; part < 2; ++part) {
      const int row = part ? bb - 1 - t : t;
      double2 nla[MB];
      if (pair) {
#pragma unroll
        for (int k = 0; k < MB; ++k)
          nla[k] = (k < ab) ? neg2(ldn(R1 + (row * ab + k) * RBLK, L)) : make_double2(0.0, 0.0);
      }
      for (int c = part; c <= row; c += 2) {
        double2 acc = cov_frag(ab + row, ab + c, true);
        if (P.want_grad) stn(Kjj + (rtri(row) + c) * RBLK, L, acc);
        if (pair) {
          double2 acc1 = make_double2(0.0, 0.0);
          const double* lc = R1 + c * ab * RBLK;
          for_desc<MB>(ab, [&](auto kc) {
            constexpr int k = decltype(kc)::value;
            mma2((k & 1) ? acc1 : acc, nla[k], ldn(lc + k * RBLK, L));
          });
          acc.x += acc1.x;
          acc.y += acc1.y;
        }
        stn(R2 + (rtri(row) + c) * RBLK, L, acc);
      }
    }
  }
  __syncthreads();
  dbg_dump(2);
  rtrace(P, s_tcur, u.uid, 3);

  // Z_i (all of it, or its first piece) travels into F while S is factored
  const int nyc = pair ? max(1, min(nyb, min(nF / ab, R_WB_BLOCKS / bb))) : max(1, min(nyb, R_WB_BLOCKS / bb));
  if (pair && tid == 0) tma_issue(stage, F, pexp + EXP_ZY, min(nyc, nyb) * ab);

  // ---- P3a: blocked Cholesky of S in place (jitchol's first, jitter-free attempt) ---------------------
  for (int J = 0; J < bb; ++J) {
    if (w == 0) {
      double av[8], wv[8];
      const int r = L.lane & 7;
      const double* src = R2 + (rtri(J) + J) * RBLK;
      if (L.lane < 8) {
#pragma unroll
        for (int v = 0; v < 8; ++v) av[v] = src[sw_off(r, v)];
      } else {
#pragma unroll
        for (int v = 0; v < 8; ++v) av[v] = (v == r) ? 1.0 : 0.0;
      }
      const int f = chol8_inv8(av, wv, L.lane);
      if (L.lane == 0 && f != 0 && *s_fail == 0) *s_fail = J * 8 + f;
      if (L.lane < 8) {
        double* dl = R2 + (rtri(J) + J) * RBLK;
        double* dw = WD + J * RBLK;
#pragma unroll
        for (int v = 0; v < 8; ++v) {
          dl[sw_off(r, v)] = av[v];
          dw[sw_off(v, r)] = wv[v];            // lane r holds column r of W_JJ = L_JJ^-1
        }
      }
    }
    __syncthreads();
    // panel: L_IJ = C_IJ W_JJ^T
    {
      const double2 bw = ldn(WD + J * RBLK, L);
      for (int I = J + 1 + w; I < bb; I += RNW) {
        double* pc = R2 + (rtri(I) + J) * RBLK;
        double2 o = make_double2(0.0, 0.0);
        mma2(o, ldn(pc, L), bw);
        stn(pc, L, o);
      }
    }
    __syncthreads();
    // trailing update: C_IK -= L_IJ L_KJ^T, J < K <= I
    {
      const int m = bb - J - 1;
      const int ntask = m * (m + 1) / 2;
      for (int t = w; t < ntask; t += RNW) {
        int ii = (int)((sqrtf(8.0f * (float)t + 1.0f) - 1.0f) * 0.5f);
        while ((ii + 1) * (ii + 2) / 2 <= t) ++ii;
        while (ii * (ii + 1) / 2 > t) --ii;
        const int kk = t - ii * (ii + 1) / 2;
        const int I = J + 1 + ii, K = J + 1 + kk;
        double* pc = R2 + (rtri(I) + K) * RBLK;
        double2 c = ldn(pc, L);
        mma2(c, neg2(ldn(R2 + (rtri(I) + J) * RBLK, L)), ldn(R2 + (rtri(K) + J) * RBLK, L));
        stn(pc, L, c);
      }
    }
    __syncthreads();
  }
  dbg_dump(3);
  rtrace(P, s_tcur, u.uid, 4);
  // log-determinant: sum of log L_tt over the unit's j rows, fixed order
  {
    double lv = 0.0;
    if (tid < bb * 8) lv = log(R2[(rtri(tid >> 3) + (tid >> 3)) * RBLK + sw_off(tid & 7, tid & 7)]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) lv += __shfl_xor_sync(0xffffffffu, lv, o);
    if (L.lane == 0) s_ld[w] = lv;
  }
  // ---- P3b: W_S = L_S^-1 in place, by descending block columns ----------------------------------------
  for (int K = bb - 1; K >= 0; --K) {
    double2 o[2];
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      o[rr] = make_double2(0.0, 0.0);
      const int I = K + 1 + w + rr * RNW;
      if (I < bb) {
        double2 acc = make_double2(0.0, 0.0);
        for (int J = K + 1; J <= I; ++J)
          mma2(acc, ldn(R2 + (rtri(I) + J) * RBLK, L), ldt(R2 + (rtri(J) + K) * RBLK, L));
        mma2(o[rr], acc, ldt(WD + K * RBLK, L));
      }
    }
    __syncthreads();
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      const int I = K + 1 + w + rr * RNW;
      if (I < bb) stn(R2 + (rtri(I) + K) * RBLK, L, neg2(o[rr]));
    }
    if (w == 0) stn(R2 + (rtri(K) + K) * RBLK, L, ldn(WD + K * RBLK, L));
    __syncthreads();
  }
  dbg_dump(4);
  rtrace(P, s_tcur, u.uid, 5);
  if (is_export) {                                  // W_b for this block's pairs
    double* dst = oexp + EXP_W;
    for (int e = tid; e < rtri(bb) * RBLK / 2; e += RNT)
      reinterpret_cast<double2*>(dst)[e] = reinterpret_cast<const double2*>(R2)[e];
  }

  // ---- P4: Z_j = W_S (Y_j - L_ji Z_i), alpha_j = W_S^T Z_j, nyc blocks of 8 outputs at a time -------------
  // tasks (row, y) of a piece are dealt round-robin to the warps; right-hand sides live in WB.
  double qsum = 0.0;
  {
    double* RB = WB;                                 // (yl * bb + row)
    bool ahead = pair;
    for (int y0 = 0; y0 < nyb; y0 += nyc) {
      const int ny = min(nyc, nyb - y0);
      const int ntask = ny * bb;
      if (pair) {
        if (!ahead) {
          if (tid == 0) tma_issue(stage, F, pexp + EXP_ZY + (long long)y0 * ab * RBLK, ny * ab);
        }
        ahead = false;
        tma_wait(stage);
      }
      for (int t = w; t < ntask; t += RNW) {
        const int yl = t / bb, row = t - yl * bb;
        const int idx = IDX[ab * 8 + row * 8 + L.g];
        const int yc = (y0 + yl) * 8 + 2 * L.q;
        double2 acc = make_double2(0.0, 0.0);
        if (idx >= 0) {
          if (yc < P.dy) acc.x = __ldg(P.Y + (long long)idx * P.dy + yc);
          if (yc + 1 < P.dy) acc.y = __ldg(P.Y + (long long)idx * P.dy + yc + 1);
        }
        if (pair) {
          const double* zi = F + yl * ab * RBLK;
          for (int k = 0; k < ab; ++k) mma2(acc, neg2(ldn(R1 + (row * ab + k) * RBLK, L)), ldt(zi + k * RBLK, L));
        }
        stn(RB + t * RBLK, L, acc);
      }
      __syncthreads();
      double2 z[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        z[i] = make_double2(0.0, 0.0);
        const int t = w + i * RNW;
        if (t < ntask) {
          const int yl = t / bb, row = t - yl * bb;
          for (int k = 0; k <= row; ++k)
            mma2(z[i], ldn(R2 + (rtri(row) + k) * RBLK, L), ldt(RB + (yl * bb + k) * RBLK, L));
          qsum += z[i].x * z[i].x + z[i].y * z[i].y;
        }
      }
      __syncthreads();
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int t = w + i * RNW;
        if (t < ntask) {
          const int yl = t / bb, row = t - yl * bb;
          stn(RB + t * RBLK, L, z[i]);
          stn(Zy + ((y0 + yl) * bb + row) * RBLK, L, z[i]);
        }
      }
      __syncthreads();
      if (P.want_grad) {
        for (int t = w; t < ntask; t += RNW) {
          const int yl = t / bb, row = t - yl * bb;
          double2 al = make_double2(0.0, 0.0);
          for (int k = row; k < bb; ++k)
            mma2(al, ldt(R2 + (rtri(k) + row) * RBLK, L), ldt(RB + (yl * bb + k) * RBLK, L));
          stn(Arow + ((long long)(ab + row) * RNYB + y0 + yl) * RBLK, L, al);
        }
      }
      __syncthreads();
    }
  }
  rtrace(P, s_tcur, u.uid, 6);
  // W_i (first piece) back into F for T while the scalars are finished
  if (pair && P.want_grad && tid == 0) tma_issue(stage, F, pexp + EXP_W, first_piece(nF, ab, tri_len));
  // |Z_j|^2 and the log-likelihood (gprf.py:542-544)
  {
    double qv = qsum;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) qv += __shfl_xor_sync(0xffffffffu, qv, o);
    if (L.lane == 0) s_q[w] = qv;
    __syncthreads();
    if (tid == 0) {
      double qt = 0.0, lt = 0.0;
      for (int i = 0; i < RNW; ++i) {
        qt += s_q[i];
        lt += s_ld[i];
      }
      double logdet = 2.0 * lt;
      if (pair) {
        qt += pexp[EXP_SCAL + 1];
        logdet += pexp[EXP_SCAL + 0];
      }
      if (is_export) {
        oexp[EXP_SCAL + 0] = logdet;
        oexp[EXP_SCAL + 1] = qt;
      }
      P.ll_u[u.uid] = -0.5 * qt - 0.5 * P.dy * logdet - 0.5 * P.dy * (double)(a + b) * 1.8378770664093454836;
      if (*s_fail != 0) {
        P.info[u.uid] = *s_fail;
        atomicOr(P.status, ST_NOTPD);
      }
    }
  }
  if (!P.want_grad) {
    __syncthreads();
    return;
  }
  asm volatile("fence.proxy.async;\n" ::: "memory");     // Z_j / alpha_j / saved K: written with ordinary stores

  // ---- P6: T = L_ji W_i (row-local, in place), V = -W_S T (in place) ------------------------------------
  if (pair) {
    bool ahead = true;
    for (int rbase = 0; rbase < bb; rbase += RNW) {
      const int row = rbase + w;
      double2 acc[MB];
#pragma unroll
      for (int c = 0; c < MB; ++c) acc[c] = make_double2(0.0, 0.0);
      staged_rows(stage, F, nF, pexp + EXP_W, ab, tri_len, tri_pos, ahead, [&](int k0, int k1, const double* base) {
        if (row < bb) {
          for (int k = k0; k < k1; ++k) {
            const double* wrow = base + (rtri(k) - rtri(k0)) * RBLK;
            const double2 av = ldn(R1 + (row * ab + k) * RBLK, L);
            for_desc<MB>(k + 1, [&](auto cc) {
              constexpr int c = decltype(cc)::value;
              mma2(acc[c], av, ldt(wrow + c * RBLK, L));
            });
          }
        }
      });
      ahead = false;
      if (row < bb) {
#pragma unroll
        for (int c = 0; c < MB; ++c)
          if (c < ab) stn(R1 + (row * ab + c) * RBLK, L, acc[c]);
      }
    }
    __syncthreads();
    // own Z_j (all of it, or its first piece) into F for alpha_i, behind the V product
    const int zrows = max(1, min(nyb, nF / bb));
    if (tid == 0) tma_issue(stage, F, Zy, zrows * bb);
    dbg_dump(6);
    rtrace(P, s_tcur, u.uid, 7);
    // V(I, .) = -sum_{k <= I} W_S(I, k) T(k, .) overwrites T(I, .), which only rows >= I read: rows
    // beyond the first 16 go first (one per warp), then rows [0, nb) with their columns split by parity
    const int nb = min(bb, RNW);
    for (int pass = (bb > RNW ? 0 : 1); pass < 2; ++pass) {
      double2 acc[MB];
#pragma unroll
      for (int c = 0; c < MB; ++c) acc[c] = make_double2(0.0, 0.0);
      if (pass == 0) {
        const int row = RNW + w;
        if (row < bb) {
          for (int k = 0; k <= row; ++k) {
            const double2 av = ldn(R2 + (rtri(row) + k) * RBLK, L);
            const double* tk = R1 + k * ab * RBLK;
            for_desc<MB>(ab, [&](auto cc) {
              constexpr int c = decltype(cc)::value;
              mma2(acc[c], av, ldt(tk + c * RBLK, L));
            });
          }
        }
        __syncthreads();
        if (row < bb) {
#pragma unroll
          for (int c = 0; c < MB; ++c)
            if (c < ab) stn(R1 + (row * ab + c) * RBLK, L, neg2(acc[c]));
        }
      } else {
        if (w < nb) {
#pragma unroll
          for (int part = 0; part < 2; ++part) {
            const int row = part ? nb - 1 - w : w;
            for (int k = 0; k <= row; ++k) {
              const double2 av = ldn(R2 + (rtri(row) + k) * RBLK, L);
              const double* tk = R1 + (k * ab + part) * RBLK;
              // columns part, part + 2, ... < ab: (ab - part + 1) / 2 of them
              for_desc<MB / 2>((ab - part + 1) >> 1, [&](auto ci) {
                constexpr int cc = decltype(ci)::value;
                mma2(acc[part * (MB / 2) + cc], av, ldt(tk + 2 * cc * RBLK, L));
              });
            }
          }
        }
        __syncthreads();
        if (w < nb) {
#pragma unroll
          for (int part = 0; part < 2; ++part) {
            const int row = part ? nb - 1 - w : w;
#pragma unroll
            for (int cc = 0; cc < MB / 2; ++cc) {
              const int c = 2 * cc + part;
              if (c < ab) stn(R1 + (row * ab + c) * RBLK, L, neg2(acc[part * (MB / 2) + cc]));
            }
          }
        }
      }
      __syncthreads();
    }
    dbg_dump(7);
    rtrace(P, s_tcur, u.uid, 8);

    // ---- P7: alpha_i = alpha_i(block) + V^T Z_j  (one row of the i part per warp, its V fragments in registers)
    ahead = true;
    for (int rbase = 0; rbase < ab; rbase += RNW) {
      const int row = rbase + w;
      double2 va[MB];
      if (row < ab) {
#pragma unroll
        for (int k = 0; k < MB; ++k)
          va[k] = (k < bb) ? ldt(R1 + (k * ab + row) * RBLK, L) : make_double2(0.0, 0.0);
      }
      staged_rows(stage, F, nF, Zy, nyb, [&](int) { return bb; }, [&](int r) { return r * bb; }, ahead,
                  [&](int y0, int y1, const double* base) {
        if (row < ab) {
          for (int yb = y0; yb < y1; ++yb) {
            const double* zj = base + (yb - y0) * bb * RBLK;
            double2 acc = ldn(pexp + EXP_AROW + ((long long)row * RNYB + yb) * RBLK, L);
            double2 acc1 = make_double2(0.0, 0.0);
            for_desc<MB>(bb, [&](auto kc) {
              constexpr int k = decltype(kc)::value;
              mma2((k & 1) ? acc1 : acc, va[k], ldt(zj + k * RBLK, L));
            });
            stn(Arow + ((long long)row * RNYB + yb) * RBLK, L, make_double2(acc.x + acc1.x, acc.y + acc1.y));
          }
        }
      });
      ahead = false;
    }
    asm volatile("fence.proxy.async;\n" ::: "memory");   // alpha_i rows
  }
  __syncthreads();
  rtrace(P, s_tcur, u.uid, 9);

  // ---- P8: G = alpha alpha^T - dy K^-1, contracted with dK in registers (gprf.py:547-584) ------------
  // Column pieces: alpha rows [c0, c1) staged in F (B operands).  Tasks = (block row r, up to GCOLS
  // columns of the piece), taken from a shared counter, biggest rows first; alpha(r) comes from L2.
  {
    const double ndy = -(double)P.dy;
    int* TT = reinterpret_cast<int*>(WB);            // task table of the piece: r | c_lo << 8 | c_hi << 16
    const int crow = max(GCOLS, (nF / RNYB) & ~(GCOLS - 1));   // alpha rows per piece (a multiple of GCOLS)
    for (int c0 = 0; c0 < nr; c0 += crow) {
      const int c1 = min(nr, c0 + crow);
      if (tid == 0) {
        tma_issue(stage, F, Arow + (long long)c0 * RNYB * RBLK, (c1 - c0) * RNYB);
        int nt = 0;
        for (int r = nr - 1; r >= c0; --r) {
          const int ce = min(c1, r + 1);
          for (int cl = c0; cl < ce; cl += GCOLS) TT[nt++] = r | (cl << 8) | (min(ce, cl + GCOLS) << 16);
        }
        *s_ntask = nt;
        *s_task = 0;
      }
      __syncthreads();
      tma_wait(stage);
      const int ntask = *s_ntask;
      while (true) {
        int t = 0;
        if (L.lane == 0) t = atomicAdd(s_task, 1);
        t = __shfl_sync(0xffffffffu, t, 0);
        if (t >= ntask) break;
        const int code = TT[t];
        const int r = code & 0xff, cl = (code >> 8) & 0xff, ch = (code >> 16) & 0xff;
        const bool irow = r < ab;
        const int wi = irow ? r : r - ab;            // row inside its part
        double2 af[RNYB];
#pragma unroll
        for (int y = 0; y < RNYB; ++y)
          af[y] = (y < nyb) ? ldn(Arow + ((long long)r * RNYB + y) * RBLK, L) : make_double2(0.0, 0.0);
        const int tr = r * 8 + L.g;
        const bool rv = IDX[tr] >= 0;
        const double* xr = xs_row(XS, tr);
        double rs[3] = {0.0, 0.0, 0.0};
        double th[MAX_NCOV];
#pragma unroll
        for (int tt = 0; tt < MAX_NCOV; ++tt) th[tt] = 0.0;
        for (int c = cl; c < ch; ++c) {
          double2 acc, kv, acc2 = make_double2(0.0, 0.0);
          if (irow) {                                 // (i row, i column): K_ii^-1 + V^T V
            acc = ldn(pexp + EXP_KINV + (long long)(rtri(wi) + c) * RBLK, L);
            kv = ldn(pexp + EXP_KSAVE + (long long)(rtri(wi) + c) * RBLK, L);
            const double* pa = R1 + wi * RBLK;
            const double* pb = R1 + c * RBLK;
            int k = 0;
            for (; k + 1 < bb; k += 2) {
              mma2(acc, ldt(pa + k * ab * RBLK, L), ldt(pb + k * ab * RBLK, L));
              mma2(acc2, ldt(pa + (k + 1) * ab * RBLK, L), ldt(pb + (k + 1) * ab * RBLK, L));
            }
            if (k < bb) mma2(acc, ldt(pa + k * ab * RBLK, L), ldt(pb + k * ab * RBLK, L));
          } else if (c < ab) {                        // (j row, i column): W_S^T V
            acc = make_double2(0.0, 0.0);
            kv = ldn(Kji + (wi * ab + c) * RBLK, L);
            int k = wi;
            for (; k + 1 < bb; k += 2) {
              mma2(acc, ldt(R2 + (rtri(k) + wi) * RBLK, L), ldt(R1 + (k * ab + c) * RBLK, L));
              mma2(acc2, ldt(R2 + (rtri(k + 1) + wi) * RBLK, L), ldt(R1 + ((k + 1) * ab + c) * RBLK, L));
            }
            if (k < bb) mma2(acc, ldt(R2 + (rtri(k) + wi) * RBLK, L), ldt(R1 + (k * ab + c) * RBLK, L));
          } else {                                    // (j row, j column): W_S^T W_S
            const int cc = c - ab;
            acc = make_double2(0.0, 0.0);
            kv = ldn(Kjj + (rtri(wi) + cc) * RBLK, L);
            int k = wi;
            for (; k + 1 < bb; k += 2) {
              mma2(acc, ldt(R2 + (rtri(k) + wi) * RBLK, L), ldt(R2 + (rtri(k) + cc) * RBLK, L));
              mma2(acc2, ldt(R2 + (rtri(k + 1) + wi) * RBLK, L), ldt(R2 + (rtri(k + 1) + cc) * RBLK, L));
            }
            if (k < bb) mma2(acc, ldt(R2 + (rtri(k) + wi) * RBLK, L), ldt(R2 + (rtri(k) + cc) * RBLK, L));
          }
          acc.x += acc2.x;
          acc.y += acc2.y;
          if (!irow && c >= ab && is_export) stn(oexp + EXP_KINV + (long long)(rtri(wi) + c - ab) * RBLK, L, acc);
          acc.x *= ndy;
          acc.y *= ndy;
          acc2 = make_double2(0.0, 0.0);
          const double* arow = F + (c - c0) * RNYB * RBLK;
          for_desc<RNYB>(nyb, [&](auto yc) {
            constexpr int y = decltype(yc)::value;
            mma2((y & 1) ? acc2 : acc, af[y], ldn(arow + y * RBLK, L));
          });
          acc.x += acc2.x;
          acc.y += acc2.y;
          // contraction of the G block with dk/dx, dk/dtheta (covariance values saved by P1 / P2 / the parent)
          double cs[2][3];
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int tc = c * 8 + 2 * L.q + e;
            const double Gv = e == 0 ? acc.x : acc.y;
            const bool cv = rv && IDX[tc] >= 0;
            if (cv && tc == tr) {
              th[0] += 0.5 * Gv;
              th[1] += 0.5 * Gv * cp.s2;
            }
            const bool off = cv && tc < tr;
            double k = e == 0 ? kv.x : kv.y;
            double gp[MAX_DX], gq[MAX_DX], gl[MAX_NLS];
            cov_grad<DFN, WFN, true>(xr, xs_row(XS, tc), cp, k, gp, gq, gl);
            const double Gm = off ? Gv : 0.0;
            th[1] += off ? Gm * k : 0.0;
#pragma unroll
            for (int d = 0; d < 3; ++d) {
              rs[d] += off ? Gm * gp[d] : 0.0;
              cs[e][d] = off ? Gm * gq[d] : 0.0;
              th[2 + d] += off ? Gm * gl[d] : 0.0;
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
              if (L.g == 0) colp[(long long)(rtri(r) + c) * COLP + (2 * L.q + e) * 3 + d] = v;
            }
        }
        // the task's row sums and theta partials
        double* tp = taskp + ((long long)r * MAXG + (cl / GCOLS)) * TASKP;
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          double v = rs[d];
          v += __shfl_xor_sync(0xffffffffu, v, 1);
          v += __shfl_xor_sync(0xffffffffu, v, 2);
          if (L.q == 0) tp[L.g * 3 + d] = v;
        }
#pragma unroll
        for (int tt = 0; tt < MAX_NCOV; ++tt) {
          double v = th[tt];
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
          if (L.lane == 0) tp[24 + tt] = v;
        }
      }
      __syncthreads();
    }
  }
  rtrace(P, s_tcur, u.uid, 10);
  // row sums + column sums, fixed order; theta
  for (int t = tid; t < nr * 8; t += RNT) {
    const int rb = t >> 3, g0 = t & 7;
    double v[3] = {0.0, 0.0, 0.0};
    for (int gi = 0; gi * GCOLS <= rb; ++gi) {
      const double* tp = taskp + ((long long)rb * MAXG + gi) * TASKP + g0 * 3;
      v[0] += tp[0];
      v[1] += tp[1];
      v[2] += tp[2];
    }
    for (int r2 = rb; r2 < nr; ++r2) {
      const double* pc = colp + (long long)(rtri(r2) + rb) * COLP + g0 * 3;
      v[0] += pc[0];
      v[1] += pc[1];
      v[2] += pc[2];
    }
    gx[t * 3] = v[0];
    gx[t * 3 + 1] = v[1];
    gx[t * 3 + 2] = v[2];
  }
  // theta: one warp per parameter; lanes stride over the task records in a fixed order, then a
  // fixed shuffle tree - the summation order does not depend on which warp ran which task
  if (w < MAX_NCOV) {
    double v = 0.0;
    for (int e = L.lane; e < nr * MAXG; e += 32) {
      const int rb = e / MAXG, gi = e - rb * MAXG;
      if (gi * GCOLS <= rb) v += taskp[(long long)e * TASKP + 24 + w];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (w == 1) v /= cp.s2;
    if (L.lane == 0) P.gth_u[(long long)u.uid * MAX_NCOV + w] = v;
  }
  __syncthreads();
  rtrace(P, s_tcur, u.uid, 11);
}

// grid: persistent CTAs (<= one per SM), RNT threads, R_SMEM_BYTES dynamic shared memory.
template <int DFN, int WFN>
__global__ void __launch_bounds__(RNT, 1) k_resident(ResParams P) {
  extern __shared__ __align__(128) double smem[];
  double* MISC = smem + R_WB_BLOCKS * RBLK;
  Stage stage;
  stage.bar = reinterpret_cast<uint64_t*>(MISC);
  stage.par = 0;
  int* s_unit = reinterpret_cast<int*>(MISC + 4);
  int* s_tcur = reinterpret_cast<int*>(MISC + 5);
  if (threadIdx.x == 0) {
    mbar_init(stage.bar, 1);
    fence_mbar_init();
    *s_tcur = 0;
  }
  __syncthreads();
  double* scratch = P.scratch + (long long)blockIdx.x * SCR_STRIDE;
  const int n_units = *P.n_order;
  while (true) {
    if (threadIdx.x == 0) *s_unit = atomicAdd(P.counter, 1);
    __syncthreads();
    const int slot = *s_unit;
    __syncthreads();
    if (slot >= n_units) break;
    Unit u;
    u.uid = P.order[slot];
    if (u.uid < P.B) {
      u.bi = -1;
      u.bj = u.uid;
    } else {
      u.bi = P.edges[2 * (u.uid - P.B)];
      u.bj = P.edges[2 * (u.uid - P.B) + 1];
    }
    u.ja = P.block_ptr[u.bj];
    u.b = (int)(P.block_ptr[u.bj + 1] - u.ja);
    u.ia = 0;
    u.a = 0;
    if (u.bi >= 0) {
      u.ia = P.block_ptr[u.bi];
      u.a = (int)(P.block_ptr[u.bi + 1] - u.ia);
    }
    if (u.bi >= 0 && u.b == 0) {
      // pair with an empty second block: the unit IS block i (computed by the block launch)
      const long long src = u.bi;
      if (threadIdx.x == 0) P.ll_u[u.uid] = P.ll_u[src];
      if (threadIdx.x < MAX_NCOV) P.gth_u[(long long)u.uid * MAX_NCOV + threadIdx.x] = P.gth_u[src * MAX_NCOV + threadIdx.x];
      for (int e = threadIdx.x; e < GX_STRIDE; e += RNT) P.gx_u[(long long)u.uid * GX_STRIDE + e] = P.gx_u[src * GX_STRIDE + e];
      continue;
    }
    u.ab = (u.a + 7) >> 3;
    u.bb = (u.b + 7) >> 3;
    const int cls = res_class(u.ab, u.bb);
    if (cls == 2) {
      if (threadIdx.x == 0) atomicOr(P.status, ST_OVERFLOW);
      continue;
    }
    if (u.b == 0) {                // empty block unit
      if (threadIdx.x == 0) {
        P.ll_u[u.uid] = 0.0;
        double* oexp = P.exports + (long long)u.bj * EXP_STRIDE;
        oexp[EXP_SCAL + 0] = 0.0;
        oexp[EXP_SCAL + 1] = 0.0;
      }
      if (threadIdx.x < MAX_NCOV) P.gth_u[(long long)u.uid * MAX_NCOV + threadIdx.x] = 0.0;
      continue;
    }
    if (u.ab <= RMAXB && u.bb <= RMAXB)
      run_unit<DFN, WFN, RMAXB>(P, u, smem, stage, scratch, cls == 1);
    else
      run_unit<DFN, WFN, BMAXB>(P, u, smem, stage, scratch, cls == 1);
  }
}

// ---------------------------------------------------------------------------------------------------
// Launch plan on the device (one CTA): which blocks have to be factored, the pair units in
// descending size order (longest first on the dynamic queue), the fit check.  No host round trip:
// block sizes never leave the GPU.
struct PlanParams {
  const long long* block_ptr;
  const int* edges;
  const unsigned char* active;     // per unit (B + E), or nullptr = all
  int B, E;
  int* order_blocks;               // out
  int* order_pairs;                // out
  int* counts;                     // out: [n_blocks, n_pairs, queue counter blocks, queue counter pairs]
  int* status;                     // out (reset here)
  int* info;                       // per unit, reset here
};

#ifndef GPRF_RES_KERNEL_ONLY
__global__ void k_res_plan(PlanParams Q) {
  extern __shared__ int sh[];          // need[B] | key[E]
  int* need = sh;
  int* key = sh + Q.B;
  const int tid = threadIdx.x, nt = blockDim.x;
  __shared__ int s_over, s_nb, s_np;
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
    key[e] = act ? (a + b) : -1;
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
  // blocks: ascending id (they are all about the same size); stable compaction by one thread per 32
  if (tid == 0) {
    int nb = 0;
    for (int bq = 0; bq < Q.B; ++bq)
      if (need[bq]) Q.order_blocks[nb++] = bq;
    s_nb = nb;
  }
  // pairs: rank sort by (size descending, id ascending)
  for (int e = tid; e < Q.E; e += nt) {
    const int ke = key[e];
    if (ke < 0) continue;
    int rank = 0;
    for (int f = 0; f < Q.E; ++f) {
      const int kf = key[f];
      rank += (kf > ke || (kf == ke && f < e)) ? 1 : 0;
    }
    Q.order_pairs[rank] = Q.B + e;
    atomicAdd(&s_np, 1);
  }
  __syncthreads();
  if (tid == 0) {
    Q.counts[0] = s_nb;
    Q.counts[1] = s_np;
    Q.counts[2] = 0;
    Q.counts[3] = 0;
    *Q.status = s_over ? ST_OVERFLOW : 0;
  }
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
        wgt = 1.0 - (double)C.deg[u];
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
    const double wgt = 1.0 - (double)C.deg[bq];
    const double* p = C.gx_u + (long long)bq * GX_STRIDE + (long long)lp * 3;
    g[0] += wgt * p[0];
    g[1] += wgt * p[1];
    g[2] += wgt * p[2];
  }
  for (int aidx = C.adj_ptr[bq]; aidx < C.adj_ptr[bq + 1]; ++aidx) {
    const int e = C.adj_edge[aidx];
    if (C.active && !C.active[C.B + e]) continue;
    int off = 0;
    if (C.adj_side[aidx]) {
      const int i = C.edges[2 * e];
      const int a = (int)(C.block_ptr[i + 1] - C.block_ptr[i]);
      off = ((a + 7) >> 3) * 8;
    }
    const double* p = C.gx_u + (long long)(C.B + e) * GX_STRIDE + (long long)(off + lp) * 3;
    g[0] += p[0];
    g[1] += p[1];
    g[2] += p[2];
  }
  const long long n = C.perm[pos];
  for (int d = 0; d < C.dx; ++d) out[1 + MAX_NCOV + n * C.dx + d] = g[d];
}
#endif  // GPRF_RES_KERNEL_ONLY

}  // namespace res
}  // namespace gprf
