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
// Coordinate record of one point as the tile kernels see it: x[0..2], 0, sin(lat), cos(lat)
// (the last two only for dfn lld: per-point terms of the haversine formula, evaluated once per
// point by prep instead of once per pair).
constexpr int XD = 6;
constexpr int MAX_NLS = 3;
constexpr int MAX_NCOV = 5;

struct CovParams {
  double nv;            // noise variance
  double s2;            // signal variance
  double il2[MAX_NLS];  // 1 / l_t^2   (0 for unused components)
  double il3[MAX_NLS];  // 1 / l_t^3
  int dx;
  int nls;
  int dfn;              // DFN_EUCLIDEAN / DFN_LLD (run-time copy for the untemplated kernels)
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
  // NaN in, NaN out (fmax above would swallow it): a NaN coordinate must poison K and fail the
  // factorisation as it does in the reference, not produce a finite objective
  return (x > 708.0) ? 0.0 : ((x != x) ? x : p * scale);
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

// sin and cos of x, branch free, for the bounded arguments of the great-circle distance (half
// differences of latitudes / longitudes in radians, latitudes; |x| up to ~1e5 is safe).
// libdevice's sincos / asin end in range-check branches and slow paths; with 16-32 pair
// evaluations per thread back to back those serialise, and the lld kernels were bound by them
// (cfg4: potrf_panel 7.7 ms against 1.4 ms for the transcendental-free trtri).  Cody-Waite
// reduction by pi/2 in two parts (fdlibm's pio2_1 / pio2_1t: n * pio2_1 is exact), fdlibm's
// __kernel_sin / __kernel_cos minimax polynomials on [-pi/4, pi/4], quadrant fix-up by selects.
// Checked against numpy on 3e6 arguments: max error 1 ulp.
__device__ __forceinline__ void sincos_bounded(double x, double& s, double& c) {
  const double MAGIC = 6755399441055744.0;
  const double kf = fma(x, 0.63661977236758134308, MAGIC);
  const int q = __double2loint(kf);
  const double n = kf - MAGIC;
  double r = fma(n, -1.57079632673412561417e+00, x);
  r = fma(n, -6.07710050650619224932e-11, r);
  const double z = r * r;
  double ps = 1.58969099521155010221e-10;
  ps = fma(ps, z, -2.50507602534068634195e-08);
  ps = fma(ps, z, 2.75573137070700676789e-06);
  ps = fma(ps, z, -1.98412698298579493134e-04);
  ps = fma(ps, z, 8.33333333332248946124e-03);
  ps = fma(ps, z, -1.66666666666666324348e-01);
  double pc = -1.13596475577881948265e-11;
  pc = fma(pc, z, 2.08757232129817482790e-09);
  pc = fma(pc, z, -2.75573143513906633035e-07);
  pc = fma(pc, z, 2.48015872894767294178e-05);
  pc = fma(pc, z, -1.38888888888741095749e-03);
  pc = fma(pc, z, 4.16666666666666019037e-02);
  const double sr = fma(r * z, ps, r);
  const double cr = fma(z, fma(z, pc, -0.5), 1.0);
  const bool sw = (q & 1) != 0;
  const double s0 = sw ? cr : sr, c0 = sw ? sr : cr;
  s = (q & 2) ? -s0 : s0;
  c = ((q + 1) & 2) ? -c0 : c0;
}

// asin(x) for x in [0, 1], branch free: fdlibm's rational approximation R(t) = p(t)/q(t) on
// t = x^2 (x < 1/2) or t = (1 - x)/2 with asin(x) = pi/2 - 2 asin(sqrt(t)) (x >= 1/2).
// Max error 2 ulp against numpy on 2e6 arguments.  x > 1 gives NaN like asin.
__device__ __forceinline__ double asin01(double x) {
  const bool big = x >= 0.5;
  const double t = big ? 0.5 * (1.0 - x) : x * x;
  const double sq = big ? sqrt(t) : x;
  double p = 3.47933107596021167570e-05;
  p = fma(p, t, 7.91534994289814532176e-04);
  p = fma(p, t, -4.00555345006794114027e-02);
  p = fma(p, t, 2.01212532134862925881e-01);
  p = fma(p, t, -3.25565818622400915405e-01);
  p = fma(p, t, 1.66666666666666657415e-01);
  p *= t;
  double q = 7.70381505559019352791e-02;
  q = fma(q, t, -6.88283971605453293030e-01);
  q = fma(q, t, 2.02094576023350569471e+00);
  q = fma(q, t, -2.40339491173441421878e+00);
  q = fma(q, t, 1.0);
  const double a = fma(sq, p / q, sq);
  return big ? fma(-2.0, a, 1.57079632679489655800e+00) + 6.12323399573676603587e-17 : a;
}

// Per-point terms of a coordinate record (prep): rec[4] = sin(lat), rec[5] = cos(lat).
__device__ __forceinline__ void point_terms(int dfn, double* rec) {
  double s = 0.0, c = 0.0;
  if (dfn == DFN_LLD) sincos_bounded(rec[1] * GPRF_DEG, s, c);
  rec[4] = s;
  rec[5] = c;
}

struct Haversine {
  double h, sp, cp, sl, cl, c1, c2, s1, s2;
};

// xp / xq are coordinate records (XD doubles, point_terms applied).
__device__ __forceinline__ Haversine haversine_terms(const double* xp, const double* xq) {
  Haversine t;
  sincos_bounded((xp[1] * GPRF_DEG - xq[1] * GPRF_DEG) * 0.5, t.sp, t.cp);
  sincos_bounded((xp[0] * GPRF_DEG - xq[0] * GPRF_DEG) * 0.5, t.sl, t.cl);
  t.s1 = xp[4]; t.c1 = xp[5];
  t.s2 = xq[4]; t.c2 = xq[5];
  t.h = t.sp * t.sp + t.c1 * t.c2 * t.sl * t.sl;
  return t;
}

// Noise-free covariance k(x_p, x_q).  xp / xq point at coordinate records (XD doubles).
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
    double d = 2.0 * GPRF_EARTH_R * asin01(sqrt(t.h));
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
    double d = 2.0 * GPRF_EARTH_R * asin01(sqrt(t.h));
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
