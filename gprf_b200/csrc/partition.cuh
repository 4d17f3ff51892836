// Point -> block assignment on the device (K8), bit-exact with the reference's
// numpy expressions.
//
//  grid  : Blocker.block_clusters (block_clustering.py:17-26):
//            argmin_b sqrt((|x|^2 * 1 - 2 x.c_b) + 1 * |c_b|^2)
//          with numpy's operation order and rounding: the squares and their sum
//          are separately rounded, the dot product is rounded the way the host
//          BLAS rounds it (dot_mode, probed by the Python layer on the actual
//          data), first index wins ties, the first NaN wins over everything
//          (np.argmin).  |c_b|^2 is computed by numpy on the host and passed in.
//  tree  : PDTree.recluster (pdtree_clustering.py:65-77) on (lon wrapped, lat):
//            a = dot(x - center, direction);  a < cut -> left, a >= cut -> right
//
// dot_mode: 0  acc = x0*c0; acc = fma(x1, c1, acc); ...      (k-ascending FMA)
//           1  acc = x_{d-1}*c_{d-1}; acc = fma(x_{d-2}, ...) (k-descending FMA)
//           2  acc = x0*c0 + x1*c1 + ...  each product and sum rounded (no FMA)
//
// Bucketing: stable radix sort of (owner, index) pairs (CUB) - ascending point
// index inside every block, as numpy's boolean-mask selection yields.
#pragma once
#include <cuda_runtime.h>
#include "resident.cuh"

namespace gprf {

template <int MODE>
__device__ __forceinline__ double dot_rounded(const double* x, const double* c, int d) {
  if (MODE == 0) {
    double acc = __dmul_rn(x[0], c[0]);
    for (int i = 1; i < d; ++i) acc = __fma_rn(x[i], c[i], acc);
    return acc;
  } else if (MODE == 1) {
    double acc = __dmul_rn(x[d - 1], c[d - 1]);
    for (int i = d - 2; i >= 0; --i) acc = __fma_rn(x[i], c[i], acc);
    return acc;
  } else {
    double acc = __dmul_rn(x[0], c[0]);
    for (int i = 1; i < d; ++i) acc = __dadd_rn(acc, __dmul_rn(x[i], c[i]));
    return acc;
  }
}

// centers: B x dx, csq: B (numpy's sum(c**2)), owner out: int32 per point.
// GT threads share a point: thread s scans centres s, s + GT, ... (each distance correctly rounded,
// independent of the others), then the partial minima are merged with numpy's sequential argmin
// rule - the smallest distance, the lowest index among equals, and the first NaN beats everything -
// which is associative, so the result is the one the sequential scan gives.
constexpr int GT = 8;
template <int MODE>
__global__ void k_assign_grid(const double* X, long long n, int dx, const double* centers, const double* csq,
                              int B, int* owner) {
  extern __shared__ double sc[];     // B * (dx + 1)
  for (int e = threadIdx.x; e < B * dx; e += blockDim.x) sc[e] = centers[e];
  for (int e = threadIdx.x; e < B; e += blockDim.x) sc[B * dx + e] = csq[e];
  __syncthreads();
  const long long p = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / GT;
  const int sub = threadIdx.x % GT;
  const bool live = p < n;
  double x[3] = {0.0, 0.0, 0.0};
  if (live)
    for (int i = 0; i < dx; ++i) x[i] = X[p * dx + i];
  double a = __dmul_rn(x[0], x[0]);
  for (int i = 1; i < dx; ++i) a = __dadd_rn(a, __dmul_rn(x[i], x[i]));
  // key = (class, distance, index): class 0 = NaN (wins, lowest index first), class 1 = number
  int best = 0x7fffffff, bcls = 2;
  double bestd = 0.0;
  for (int b = sub; b < B; b += GT) {
    const double dot = dot_rounded<MODE>(x, sc + b * dx, dx);
    const double t = __dadd_rn(__dsub_rn(a, 2.0 * dot), sc[B * dx + b]);
    const double d = __dsqrt_rn(t);
    const int cls = (d != d) ? 0 : 1;
    const bool better = cls < bcls || (cls == bcls && cls == 1 && d < bestd);   // ascending b: ties keep the earlier
    if (better) {
      bcls = cls;
      bestd = d;
      best = b;
    }
  }
#pragma unroll
  for (int o = 1; o < GT; o <<= 1) {
    const int oc = __shfl_xor_sync(0xffffffffu, bcls, o);
    const double od = __shfl_xor_sync(0xffffffffu, bestd, o);
    const int ob = __shfl_xor_sync(0xffffffffu, best, o);
    const bool better = oc < bcls || (oc == bcls && (oc == 1 ? (od < bestd || (od == bestd && ob < best)) : ob < best));
    if (better) {
      bcls = oc;
      bestd = od;
      best = ob;
    }
  }
  if (live && sub == 0) owner[p] = best;
}

struct TreeParams {
  const double* center;      // nnodes x 2
  const double* direction;   // nnodes x 2
  const double* cut;         // nnodes
  const int* child;          // nnodes x 2; negative = -(leaf id) - 1
  int root;
  double wrap_add, wrap_mod; // lon -> (lon + wrap_add) % wrap_mod - wrap_add   (22, 360)
};

template <int MODE>
__global__ void k_assign_tree(const double* X, long long n, int dx, TreeParams Tp, int* owner) {
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  double q[2];
  {
    // numpy float remainder: fmod, then shift into the sign of the divisor
    double r = fmod(__dadd_rn(X[p * dx], Tp.wrap_add), Tp.wrap_mod);
    if (r != 0.0) {
      if (r < 0.0) r = __dadd_rn(r, Tp.wrap_mod);
    } else {
      r = 0.0;
    }
    q[0] = __dsub_rn(r, Tp.wrap_add);
    q[1] = X[p * dx + 1];
  }
  int node = Tp.root;
  int guard = 0;
  while (node >= 0 && guard++ < 64) {
    double v[2] = {__dsub_rn(q[0], Tp.center[2 * node]), __dsub_rn(q[1], Tp.center[2 * node + 1])};
    const double a = dot_rounded<MODE>(v, Tp.direction + 2 * node, 2);
    node = (a < Tp.cut[node]) ? Tp.child[2 * node] : Tp.child[2 * node + 1];
  }
  owner[p] = -node - 1;
}

__global__ void k_iota(int* idx, long long n) {
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p < n) idx[p] = (int)p;
}

// sorted owner keys -> block_ptr[b] = first position with key >= b; perm64 from the sorted indices
__global__ void k_block_bounds(const int* keys, const int* idx_sorted, long long n, int B, long long* block_ptr,
                               long long* perm64) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) perm64[t] = idx_sorted[t];
  if (t <= B) {
    long long lo = 0, hi = n;
    while (lo < hi) {
      long long mid = (lo + hi) >> 1;
      if (keys[mid] < (int)t) lo = mid + 1; else hi = mid;
    }
    block_ptr[t] = lo;
  }
}

// Small problems (n <= 65535 * BK_WARPS is not needed: n < 2^21, B <= BK_MAXB): ONE CTA buckets the
// points by owner - a stable counting sort, so that every block lists its points in ascending order
// like numpy's nonzero() - and leaves perm / pos_block / block_ptr exactly as the radix-sort path
// does.  Replaces three cub launches + k_block_bounds (27 us at n = 10^4) by one of a few us; with
// `plan` set it goes on to build the resident path's launch plan (resident.cuh) in the same launch.
// Warp w owns the contiguous point range [w L, (w + 1) L): (1) per-warp histograms, (2) per block a
// prefix over the warps, (3) every warp places its points in order, ranks inside a group of 32 by
// __match_any_sync.  Shared memory: BK_WARPS x B ints + (B + 1) ints (+ B + E ints for the plan).
constexpr int BK_WARPS = 32;
constexpr int BK_MAXB = 1024;
constexpr long long BK_MAXN = 1 << 18;
struct BucketPlan {
  int enabled;
  res::PlanParams Q;
};
// Launch geometry: ONE CTA does everything (gridDim.x == 1), or 1 + npl CTAs split the work - CTA 0 counts
// the blocks' totals, writes block_ptr and builds the launch plan (which needs the block SIZES only), while
// CTAs 1 .. npl place the points: the point range is cut into BK_WARPS * npl sub-ranges, every placement CTA
// builds the histograms of ALL sub-ranges (redundantly; 10^4 shared-memory atomics) and places its own
// BK_WARPS of them.  Same permutation by construction (sub-ranges in order, points of a sub-range in order);
// the placement pass (6.6 us as one chain of n / 32 / BK_WARPS steps per warp) and the plan (~7 us) then
// overlap instead of following each other.
__device__ __forceinline__ void bucket_starts(int* start, int B) {     // block totals -> block starts (warp 0)
  const int lane = threadIdx.x & 31;
  const int per = (B + 31) / 32;
  int tot = 0;
  for (int b = lane * per; b < min(B, (lane + 1) * per); ++b) tot += start[b];
  int pre = tot;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, pre, o);
    if (lane >= o) pre += v;
  }
  int at = pre - tot;
  for (int b = lane * per; b < min(B, (lane + 1) * per); ++b) {
    const int c = start[b];
    start[b] = at;
    at += c;
  }
  if (lane == 31) start[B] = pre;
}

__global__ void __launch_bounds__(BK_WARPS * 32, 1)
k_bucket_small(const int* owner, long long n, int B, long long* block_ptr, long long* perm64, int* pos_block,
               BucketPlan plan) {
  extern __shared__ int bk_sh[];
  const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
  const bool split = gridDim.x > 1;
  if (split && blockIdx.x == 0) {
    // totals + plan
    int* start = bk_sh;                  // [B + 1]
    int* plan_sh = start + B + 1;        // (B + E) ints + the plan's histograms
    for (int b = tid; b <= B; b += blockDim.x) start[b] = 0;
    __syncthreads();
    for (long long p = tid; p < n; p += blockDim.x) atomicAdd(&start[owner[p]], 1);
    __syncthreads();
    if (w == 0) bucket_starts(start, B);
    __syncthreads();
    for (int b = tid; b <= B; b += blockDim.x) block_ptr[b] = start[b];
    if (plan.enabled) {
      __threadfence_block();
      __syncthreads();                   // block_ptr (global, this CTA's own writes) is read back by the plan
      res::res_plan_body(plan.Q, plan_sh);
    }
    return;
  }
  const int npl = split ? (int)gridDim.x - 1 : 1;          // placement CTAs
  const int cta = split ? (int)blockIdx.x - 1 : 0;
  const int R = BK_WARPS * npl;                             // sub-ranges
  int* wh = bk_sh;                       // [R][B]: counts, then exclusive prefixes over the sub-ranges
  int* start = bk_sh + R * B;            // [B + 1]: block totals, then block starts
  int* plan_sh = start + B + 1;          // (B + E) ints for the plan (single-CTA launch)
  for (int e = tid; e < R * B; e += blockDim.x) wh[e] = 0;
  __syncthreads();
  const long long L = (n + R - 1) / R;
  const int mine = cta * BK_WARPS + w;   // the sub-range this warp places
  const long long p0 = mine * L < n ? mine * L : n, p1 = p0 + L < n ? p0 + L : n;
  // the warp's owners: the first BK_PRE rounds stay in registers for the placement pass
  constexpr int BK_PRE = 16;
  int own[BK_PRE];
#pragma unroll
  for (int k = 0; k < BK_PRE; ++k) {
    const long long p = p0 + lane + 32 * k;
    own[k] = p < p1 ? owner[p] : -1;
  }
#pragma unroll
  for (int k = 0; k < BK_PRE; ++k)
    if (own[k] >= 0) atomicAdd(&wh[mine * B + own[k]], 1);
  for (long long p = p0 + lane + 32 * BK_PRE; p < p1; p += 32) atomicAdd(&wh[mine * B + owner[p]], 1);
  // (split launch) the other placement CTAs' sub-ranges: every CTA needs all the histograms
  for (int r = w; r < R; r += BK_WARPS) {
    if (r == mine || r / BK_WARPS == cta) continue;
    const long long q0 = r * L < n ? r * L : n, q1 = q0 + L < n ? q0 + L : n;
    for (long long p = q0 + lane; p < q1; p += 32) atomicAdd(&wh[r * B + owner[p]], 1);
  }
  __syncthreads();
  // per block: exclusive prefix over the sub-ranges (in place) and the block total
  for (int b = tid; b < B; b += blockDim.x) {
    int c = 0;
#pragma unroll 8
    for (int v = 0; v < R; ++v) {
      const int x = wh[v * B + b];
      wh[v * B + b] = c;
      c += x;
    }
    start[b] = c;
  }
  __syncthreads();
  if (w == 0) bucket_starts(start, B);
  __syncthreads();
  if (!split)
    for (int b = tid; b <= B; b += blockDim.x) block_ptr[b] = start[b];
  // placement: every warp walks its points in order; rank inside a group of 32 by __match_any_sync
  int kbits = 1;
  while ((1 << kbits) < B) ++kbits;
  auto place = [&](long long q0, int o, bool live) {
    // lanes holding the same owner: one ballot per key bit (__match_any_sync costs ~1500 cycles per call
    // with ~27 distinct keys among the 32 lanes: measured 19 000 -> 13 000 cycles for this pass)
    unsigned grp = __ballot_sync(0xffffffffu, live);
    for (int bit = 0; bit < kbits; ++bit) {
      const bool one = (o >> bit) & 1;
      const unsigned mb = __ballot_sync(0xffffffffu, one);
      grp &= one ? mb : ~mb;
    }
    if (!live) grp = 1u << lane;
    const int rank = __popc(grp & ((1u << lane) - 1u));
    const int leader = __ffs(grp) - 1;
    int base = 0;
    if (live && lane == leader) {
      base = wh[mine * B + o];
      wh[mine * B + o] = base + __popc(grp);
    }
    base = __shfl_sync(0xffffffffu, base, leader);
    if (live) {
      const int pos = start[o] + base + rank;
      perm64[pos] = q0 + lane;
      pos_block[pos] = o;
    }
    __syncwarp();
  };
#pragma unroll
  for (int k = 0; k < BK_PRE; ++k) {
    const long long q0 = p0 + 32 * k;
    if (q0 < p1) place(q0, own[k], own[k] >= 0);
  }
  for (long long q0 = p0 + 32 * BK_PRE; q0 < p1; q0 += 32) {
    const long long p = q0 + lane;
    const bool live = p < p1;
    place(q0, live ? owner[p] : 0, live);
  }
  if (!split && plan.enabled) {
    __threadfence_block();
    __syncthreads();                     // block_ptr (global, this CTA's own writes) is read back by the plan
    res::res_plan_body(plan.Q, plan_sh);
  }
}

}  // namespace gprf
