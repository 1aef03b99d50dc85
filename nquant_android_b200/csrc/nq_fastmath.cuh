// nq_fastmath.cuh -- plain-double evaluation of the H' and R_T pieces of CIEDE2000 with an error
// budget, used as a filter in front of the correctly rounded double-double kernels of nq_math.h.
//
// The reference narrows both pieces to float (CL:185, CL:193). The kernels below approximate the same
// real-valued expressions to ~1e-13; whenever the interval [x - e, x + e] around the approximation
// rounds to ONE float, that float is the value the correctly rounded path produces too (float rounding
// is monotone). Otherwise -- and whenever a branch of the reference's hue logic (CL:139-176) sits
// within a guard band of its threshold -- the caller runs the exact path. The filter therefore never
// changes a result; it only decides how much work a candidate costs. tests/test_gpu_parity.py
// compares millions of pairs against the oracle bit for bit, including constructed near-threshold ones.
#pragma once
#include "nq_math.h"
#include "nq_color.h"

namespace nqf {

// atan2 for finite x, y with x != 0 and y != 0. |result - atan2(y, x)| < 1e-15.
__device__ __forceinline__ double atan2_fast(double y, double x) {
  const double ay = fabs(y), ax = fabs(x);
  const bool swap = ay > ax;
  const double a = swap ? ax : ay, b = swap ? ay : ax;              // t = a / b in (0, 1]
  int i = __float2int_rn(__fdividef((float)a, (float)b) * 16.f);    // nearest table node, any estimate will do
  i = min(max(i, 0), 16);
  const double c = (double)i * 0.0625;
  // atan(t) = atan(c) + atan(u), u = (t - c) / (1 + t c) = (a - c b) / (b + c a), |u| < 1/30
  const double u = __fma_rn(-c, b, a) / __fma_rn(c, a, b);
  const double z = u * u;
  double p = 1.0 / 13.0;
  p = __fma_rn(p, z, -1.0 / 11.0);
  p = __fma_rn(p, z, 1.0 / 9.0);
  p = __fma_rn(p, z, -1.0 / 7.0);
  p = __fma_rn(p, z, 1.0 / 5.0);
  p = __fma_rn(p, z, -1.0 / 3.0);
  double at = nqm::atan_16_tab(i, 0) + (__fma_rn(u * z, p, u) + nqm::atan_16_tab(i, 1));
  if (swap) at = (nqm::PIO2_HI - at) + nqm::PIO2_LO;
  if (x < 0.0) at = (nqm::PI_HI - at) + nqm::PI_LO;
  return y < 0.0 ? -at : at;
}

// x = n pi/2 + r, |x| < 64
__device__ __forceinline__ double reduce_pio2(double x, int* q) {
  const double fn = nqm::rint_(x * nqm::TWO_OVER_PI);
  double r = __fma_rn(-fn, nqm::PIO2_1, x);
  r = __fma_rn(-fn, nqm::PIO2_2, r);
  r = __fma_rn(-fn, nqm::PIO2_3, r);
  *q = (int)fn & 3;
  return r;
}
__device__ __forceinline__ double sin_poly(double r) {   // |r| <= pi/4 + slack
  const double z = r * r;
  double p = 1.0 / 355687428096000.0;                     // 1/17!
  p = __fma_rn(p, z, -1.0 / 1307674368000.0);             // 1/15!
  p = __fma_rn(p, z, 1.0 / 6227020800.0);                 // 1/13!
  p = __fma_rn(p, z, -1.0 / 39916800.0);                  // 1/11!
  p = __fma_rn(p, z, 1.0 / 362880.0);                     // 1/9!
  p = __fma_rn(p, z, -1.0 / 5040.0);                      // 1/7!
  p = __fma_rn(p, z, 1.0 / 120.0);                        // 1/5!
  p = __fma_rn(p, z, -1.0 / 6.0);
  return __fma_rn(r * z, p, r);
}
__device__ __forceinline__ double cos_poly(double r) {
  const double z = r * r;
  double p = 1.0 / 20922789888000.0;                      // 1/16!
  p = __fma_rn(p, z, -1.0 / 87178291200.0);               // 1/14!
  p = __fma_rn(p, z, 1.0 / 479001600.0);                  // 1/12!
  p = __fma_rn(p, z, -1.0 / 3628800.0);                   // 1/10!
  p = __fma_rn(p, z, 1.0 / 40320.0);                      // 1/8!
  p = __fma_rn(p, z, -1.0 / 720.0);                       // 1/6!
  p = __fma_rn(p, z, 1.0 / 24.0);                         // 1/4!
  p = __fma_rn(p, z, -0.5);
  return __fma_rn(z, p, 1.0);
}
// |error| < 5e-16 for |x| < 64
__device__ __forceinline__ double sin_fast(double x) {
  int q;
  const double r = reduce_pio2(x, &q);
  const double v = (q & 1) ? cos_poly(r) : sin_poly(r);
  return (q & 2) ? -v : v;
}
__device__ __forceinline__ double cos_fast(double x) {
  int q;
  const double r = reduce_pio2(x, &q);
  const double v = (q & 1) ? sin_poly(r) : cos_poly(r);
  return ((q + 1) & 2) ? -v : v;
}
// exp(x) for -600 < x <= 0, relative error < 5e-16
__device__ __forceinline__ double exp_fast(double x) {
  const double fk = nqm::rint_(x * nqm::INV_LN2_32);
  const int kk = (int)fk;
  double r = __fma_rn(-fk, nqm::LN2_32_HI, x);
  r = __fma_rn(-fk, nqm::LN2_32_LO, r);
  double p = 1.0 / 720.0;
  p = __fma_rn(p, r, 1.0 / 120.0);
  p = __fma_rn(p, r, 1.0 / 24.0);
  p = __fma_rn(p, r, 1.0 / 6.0);
  p = __fma_rn(p, r, 0.5);
  p = __fma_rn(p, r, 1.0);
  p = p * r;                                              // expm1(r), |r| <= 0.011
  const int j = kk & 31;
  const double T = nqm::exp2_32_tab(j, 0);
  return (__fma_rn(T, p, nqm::exp2_32_tab(j, 1)) + T) * nqm::pow2i((kk - j) / 32);
}

// H' and R_T terms of CIEDE2000 as the floats CIELABConvertor.H_prime_div_k_L_S_L (CL:120-185) and
// R_T (CL:187-194) return, or false when the exact path has to decide.
__device__ __forceinline__ bool ciede_HRT_fast(float B1, float B2, const nq::CiedeC& c, float tC, float* tH, float* tRT) {
  const double b1 = (double)B1, b2 = (double)B2;
  if (b1 == 0.0 || b2 == 0.0 || c.a1p == 0.0 || c.a2p == 0.0) return false;     // axis cases
  const double deg360 = (double)nq::deg2rad(360.f), deg180 = (double)nq::deg2rad(180.f);
  const double GUARD = 1e-12;
  double h1 = atan2_fast(b1, c.a1p), h2 = atan2_fast(b2, c.a2p);    // |error| < 1e-15 each
  if (b1 < 0.0) h1 += deg360;                                        // sign(atan2) == sign(y)
  if (b2 < 0.0) h2 += deg360;
  double dh = h2 - h1;
  const double hsum = h1 + h2;
  if (fabs(fabs(dh) - deg180) < GUARD || fabs(hsum - deg360) < GUARD) return false;
  const bool wide = fabs(dh) > deg180;
  if (dh < -deg180) dh += deg360;
  else if (dh > deg180) dh -= deg360;
  const double sq = nqm::sqrt_(c.C1p * c.C2p);
  const double dHP = 2.0 * sq * sin_fast(dh / 2.0);                  // |error| <= sq * 1.5e-14
  double bh;
  if (!wide) bh = hsum / 2.0;
  else if (hsum < deg360) bh = (hsum + deg360) / 2.0;
  else bh = (hsum - deg360) / 2.0;                                   // |error| < 1e-14
  const double barC = (c.C1p + c.C2p) / 2.0;
  const double T = 1.0 - (0.17 * cos_fast(bh - (double)nq::deg2rad(30.f))) + (0.24 * cos_fast(2.0 * bh)) +
                   (0.32 * cos_fast((3.0 * bh) + (double)nq::deg2rad(6.f))) - (0.20 * cos_fast((4.0 * bh) - (double)nq::deg2rad(63.f)));
  const double SH = 1 + ((double)0.015f * barC * T);                // >= 1, relative error < 1e-13
  const double xH = dHP / SH;
  const double eH = sq * 3e-14 + fabs(xH) * 2e-13;
  const float hl = (float)(xH - eH), hh = (float)(xH + eH);
  if (hl != hh) return false;
  const double q = (bh - (double)nq::deg2rad(275.f)) / (double)nq::deg2rad(25.f);
  const double dT = (double)nq::deg2rad(30.f) * exp_fast(-(q * q)); // relative error < 1e-12
  const double c2 = barC * barC, c4 = c2 * c2;
  const double p7 = (c4 * c2) * barC;
  const double RC = 2.0 * nqm::sqrt_(p7 / (p7 + 6103515625.0));
  const double xR = ((-sin_fast(2.0 * dT)) * RC) * (double)tC * (double)hl;
  const double eR = fabs(xR) * 4e-12;
  const float rl = (float)(xR - eR), rh = (float)(xR + eR);
  if (rl != rh) return false;
  *tH = hl; *tRT = rl;
  return true;
}

}  // namespace nqf
