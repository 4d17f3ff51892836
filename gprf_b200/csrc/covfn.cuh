// Covariance-function arithmetic evaluated in registers (K0 / K5 / K6).
//
// Replaces treegp's VectorTree::kernel_matrix / kernel_deriv_wrt_xi_row /
// kernel_deriv_wrt_i as called from gprf.py:339,342,353,373-374 (treegp is an
// external dependency of the reference, pinned by its README.md:4; the
// formulas are the published definitions, restated in SURVEY.md section 8a).
//
//   dfn euclidean : r^2 = sum_i (dx_i / l_i)^2
//   dfn lld       : x = (lon deg, lat deg, depth); d = haversine km (R = 6371);
//                   r^2 = (d/l0)^2 + (dz/l1)^2          (run_seismic.py:19-63,230-233)
//   wfn se        : w = s2 exp(-r^2)
//   wfn matern32  : w = s2 (1 + sqrt3 r) exp(-sqrt3 r)
//
// Derivatives are expressed through wr = w'(r)/r, which is finite at r = 0:
//   se: wr = -2 w          matern32: wr = -3 s2 exp(-sqrt3 r)
//   dk/dx_{p,i} = wr * (r dr/dx_{p,i}),   dk/dl_t = -wr * c_t^2 / l_t^3
#pragma once
#include <math.h>

namespace gprf {

enum { DFN_EUCLIDEAN = 0, DFN_LLD = 1 };
enum { WFN_SE = 0, WFN_MATERN32 = 1 };

constexpr int MAX_DX = 3;
constexpr int MAX_NLS = 3;
constexpr int MAX_NCOV = 5;

struct CovParams {
  double nv;            // noise variance
  double s2;            // signal variance
  double il2[MAX_NLS];  // 1 / l_t^2   (0 for unused components)
  double il3[MAX_NLS];  // 1 / l_t^3
  int dx;
  int nls;
};

#define GPRF_EARTH_R 6371.0
#define GPRF_DEG 0.017453292519943295769  // pi / 180
#define GPRF_SQRT3 1.7320508075688772935

// exp(-x) for x >= 0, branch free.  libdevice's exp() ends in a range-check branch, which keeps
// the scheduler from interleaving independent evaluations: 160 cycles each however many are in
// flight, against ~45 for this one at ILP 4 (scripts/fp64_latency.cu).  The tile kernels evaluate
// 16-32 covariances per thread back to back, so this is what bounds their prologues.
// Cody-Waite reduction x = n ln2 + r, |r| <= ln2/2, degree-13 Taylor polynomial (truncation
// 4e-18), 2^n through the exponent field.  Max error ~1.5 ulp; results below 2^-1021 (x > 708)
// are flushed to 0 instead of going subnormal.
__device__ __forceinline__ double exp_neg(double x) {
  const double t = fmax(-x, -708.0);
  const double MAGIC = 6755399441055744.0;               // 2^52 + 2^51: rounds to nearest integer
  const double kf = fma(t, 1.4426950408889634074, MAGIC);
  const int ni = __double2loint(kf);
  const double n = kf - MAGIC;
  double r = fma(n, -6.93147180369123816490e-01, t);     // ln2 high part (fdlibm split)
  r = fma(n, -1.90821492927058770002e-10, r);            // ln2 low part
  double p = 1.6059043836821613e-10;                     // 1/13!
  p = fma(p, r, 2.08767569878681e-09);                   // 1/12!
  p = fma(p, r, 2.505210838544172e-08);                  // 1/11!
  p = fma(p, r, 2.755731922398589e-07);                  // 1/10!
  p = fma(p, r, 2.7557319223985893e-06);                 // 1/9!
  p = fma(p, r, 2.48015873015873e-05);                   // 1/8!
  p = fma(p, r, 1.984126984126984e-04);                  // 1/7!
  p = fma(p, r, 1.388888888888889e-03);                  // 1/6!
  p = fma(p, r, 8.333333333333333e-03);                  // 1/5!
  p = fma(p, r, 4.1666666666666664e-02);                 // 1/4!
  p = fma(p, r, 1.6666666666666666e-01);                 // 1/3!
  p = fma(p, r, 0.5);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
  const double scale = __hiloint2double((ni + 1023) << 20, 0);
  return (x > 708.0) ? 0.0 : p * scale;
}

template <int WFN>
__device__ __forceinline__ double weight_from_r2(double r2, double s2) {
  if (WFN == WFN_SE) {
    return s2 * exp_neg(r2);
  } else {
    double a = GPRF_SQRT3 * sqrt(r2);
    return s2 * (1.0 + a) * exp_neg(a);
  }
}

// w and wr = w'(r)/r
template <int WFN>
__device__ __forceinline__ void weight_and_wr(double r2, double s2, double& w, double& wr) {
  if (WFN == WFN_SE) {
    w = s2 * exp_neg(r2);
    wr = -2.0 * w;
  } else {
    double a = GPRF_SQRT3 * sqrt(r2);
    double e = s2 * exp_neg(a);
    w = (1.0 + a) * e;
    wr = -3.0 * e;
  }
}

struct Haversine {
  double h, sp, cp, sl, cl, c1, c2, s1, s2;
};

__device__ __forceinline__ Haversine haversine_terms(const double* xp, const double* xq) {
  Haversine t;
  double p1 = xp[1] * GPRF_DEG, p2 = xq[1] * GPRF_DEG;
  sincos((p1 - p2) * 0.5, &t.sp, &t.cp);
  sincos((xp[0] * GPRF_DEG - xq[0] * GPRF_DEG) * 0.5, &t.sl, &t.cl);
  sincos(p1, &t.s1, &t.c1);
  sincos(p2, &t.s2, &t.c2);
  t.h = t.sp * t.sp + t.c1 * t.c2 * t.sl * t.sl;
  return t;
}

// Noise-free covariance k(x_p, x_q).  xp / xq point at MAX_DX+1 doubles.
template <int DFN, int WFN>
__device__ __forceinline__ double cov_value(const double* xp, const double* xq, const CovParams& cp) {
  double r2;
  if (DFN == DFN_EUCLIDEAN) {
    r2 = 0.0;
#pragma unroll
    for (int i = 0; i < MAX_DX; ++i) {
      double d = xp[i] - xq[i];
      r2 += d * d * cp.il2[i];
    }
  } else {
    Haversine t = haversine_terms(xp, xq);
    double d = 2.0 * GPRF_EARTH_R * asin(sqrt(t.h));
    double dz = xp[2] - xq[2];
    r2 = d * d * cp.il2[0] + dz * dz * cp.il2[1];
  }
  return weight_from_r2<WFN>(r2, cp.s2);
}

// Covariance value and all first derivatives for one ordered pair (p, q), p != q:
//   gp[i] = dk/dx_{p,i},  gq[i] = dk/dx_{q,i},  gl[t] = dk/dl_t
// HAVE_K: k holds the covariance value on entry (saved by the factorisation, which had to
// evaluate it anyway), so the exponential is not recomputed:
//   se: wr = -2 k          matern32: wr = -3 k / (1 + sqrt3 r)
template <int DFN, int WFN, bool HAVE_K>
__device__ __forceinline__ void cov_grad(const double* xp, const double* xq, const CovParams& cp,
                                         double& k, double gp[MAX_DX], double gq[MAX_DX],
                                         double gl[MAX_NLS]) {
  double wr;
  if (DFN == DFN_EUCLIDEAN) {
    double d[MAX_DX];
    double r2 = 0.0;
#pragma unroll
    for (int i = 0; i < MAX_DX; ++i) {
      d[i] = xp[i] - xq[i];
      r2 += d[i] * d[i] * cp.il2[i];
    }
    if (HAVE_K) {
      wr = (WFN == WFN_SE) ? -2.0 * k : -3.0 * k / (1.0 + GPRF_SQRT3 * sqrt(r2));
    } else {
      weight_and_wr<WFN>(r2, cp.s2, k, wr);
    }
#pragma unroll
    for (int i = 0; i < MAX_DX; ++i) {
      double g = wr * d[i] * cp.il2[i];
      gp[i] = g;
      gq[i] = -g;
      gl[i] = -wr * d[i] * d[i] * cp.il3[i];
    }
  } else {
    Haversine t = haversine_terms(xp, xq);
    double d = 2.0 * GPRF_EARTH_R * asin(sqrt(t.h));
    double dz = xp[2] - xq[2];
    double r2 = d * d * cp.il2[0] + dz * dz * cp.il2[1];
    if (HAVE_K) {
      wr = (WFN == WFN_SE) ? -2.0 * k : -3.0 * k / (1.0 + GPRF_SQRT3 * sqrt(r2));
    } else {
      weight_and_wr<WFN>(r2, cp.s2, k, wr);
    }
    // d * dd/dh / l0^2 ; the reference's 0 * inf at h == 0 is mapped to 0
    double pref = 0.0;
    if (t.h > 0.0) pref = wr * d * (GPRF_EARTH_R / sqrt(t.h * (1.0 - t.h))) * cp.il2[0];
    if (!isfinite(pref)) pref = 0.0;
    double slcl = t.c1 * t.c2 * t.sl * t.cl * GPRF_DEG;
    double sl2 = t.sl * t.sl;
    gp[0] = pref * slcl;
    gq[0] = -pref * slcl;
    gp[1] = pref * (t.sp * t.cp - t.s1 * t.c2 * sl2) * GPRF_DEG;
    gq[1] = pref * (-t.sp * t.cp - t.s2 * t.c1 * sl2) * GPRF_DEG;
    double gz = wr * dz * cp.il2[1];
    gp[2] = gz;
    gq[2] = -gz;
    gl[0] = -wr * d * d * cp.il3[0];
    gl[1] = -wr * dz * dz * cp.il3[1];
    gl[2] = 0.0;
  }
}

}  // namespace gprf
