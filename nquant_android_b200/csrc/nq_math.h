// nq_math.h -- one FMA-explicit, contraction-free implementation of the transcendental functions
// the quantizer path needs (pow, exp, tanh, cbrt, atan2, sin, cos), compiled unchanged for the host
// (g++ -ffp-contract=off) and for sm_100a (nvcc -fmad=false).
//
// Why this exists: the reference evaluates java.lang.Math.{pow,exp,tanh,cbrt,atan2,sin,cos}
// (CIELABConvertor.java:74,96,107,130,162,179-192; GilbertCurve.java:102,118-119,255,341;
// PnnLABQuantizer.java:62,121,225-241,348; androidx ColorUtils pow calls behind CL:61,78).
// Java only promises <=1-2 ulp for these, so "the" reference value is platform dependent. We pin
// one answer: double-double kernels whose result is the correctly rounded value except in
// astronomically rare hard cases (error < 0.5 + 2^-9 ulp), built only from IEEE +,-,*,/,sqrt,fma,
// which round identically on x86-64 and on the GPU. Host and device therefore agree bit for bit.
#pragma once
#include <stdint.h>
#include <string.h>
#include <math.h>

#if defined(__CUDACC__)
#define NQ_HD __host__ __device__ __forceinline__
#else
#define NQ_HD inline
#endif
#define NQ_CONST static constexpr

#include "nq_math_tables.h"

namespace nqm {

static const double EXP2_32_H[32][2] = NQ_EXP2_32_INIT;
static const double ATAN_16_H[17][2] = NQ_ATAN_16_INIT;
#if defined(__CUDACC__)
static __device__ const double EXP2_32_D[32][2] = NQ_EXP2_32_INIT;
static __device__ const double ATAN_16_D[17][2] = NQ_ATAN_16_INIT;
#endif

NQ_HD double exp2_32_tab(int j, int k) {
#if defined(__CUDA_ARCH__)
  return EXP2_32_D[j][k];
#else
  return EXP2_32_H[j][k];
#endif
}
NQ_HD double atan_16_tab(int j, int k) {
#if defined(__CUDA_ARCH__)
  return ATAN_16_D[j][k];
#else
  return ATAN_16_H[j][k];
#endif
}

NQ_HD double fma_(double a, double b, double c) {
#if defined(__CUDA_ARCH__)
  return __fma_rn(a, b, c);
#else
  return __builtin_fma(a, b, c);
#endif
}
NQ_HD double sqrt_(double a) {
#if defined(__CUDA_ARCH__)
  return __dsqrt_rn(a);
#else
  return __builtin_sqrt(a);
#endif
}
NQ_HD uint64_t d2bits(double x) {
#if defined(__CUDA_ARCH__)
  return (uint64_t)__double_as_longlong(x);
#else
  uint64_t u; memcpy(&u, &x, 8); return u;
#endif
}
NQ_HD double bits2d(uint64_t u) {
#if defined(__CUDA_ARCH__)
  return __longlong_as_double((long long)u);
#else
  double x; memcpy(&x, &u, 8); return x;
#endif
}
NQ_HD double fabs_(double x) { return bits2d(d2bits(x) & 0x7fffffffffffffffULL); }
NQ_HD bool isnan_(double x) { return x != x; }
// round to nearest integer, ties to even; |x| < 2^51
NQ_HD double rint_(double x) {
  const double M = 6755399441055744.0;  // 1.5 * 2^52
  return (x + M) - M;
}
// 2^k for -1022 <= k <= 1023
NQ_HD double pow2i(int k) { return bits2d((uint64_t)(k + 1023) << 52); }
NQ_HD double scalbn_(double x, int k) {
  if (k > 1023) { x *= pow2i(1023); k -= 1023; if (k > 1023) k = 1023; }
  else if (k < -1022) { x *= pow2i(-1022); k += 1022; if (k < -1022) k = -1022; }
  return x * pow2i(k);
}

struct dd { double hi, lo; };

NQ_HD dd two_sum(double a, double b) {
  double s = a + b, bb = s - a;
  double e = (a - (s - bb)) + (b - bb);
  return dd{s, e};
}
NQ_HD dd fast_two_sum(double a, double b) {  // needs |a| >= |b| or a == 0
  double s = a + b;
  return dd{s, b - (s - a)};
}
NQ_HD dd two_prod(double a, double b) {
  double p = a * b;
  return dd{p, fma_(a, b, -p)};
}
NQ_HD dd dd_add(dd a, dd b) {
  dd s = two_sum(a.hi, b.hi), t = two_sum(a.lo, b.lo);
  double c = s.lo + t.hi;
  dd v = fast_two_sum(s.hi, c);
  double w = t.lo + v.lo;
  return fast_two_sum(v.hi, w);
}
NQ_HD dd dd_add_d(dd a, double b) {
  dd s = two_sum(a.hi, b);
  return fast_two_sum(s.hi, s.lo + a.lo);
}
NQ_HD dd dd_neg(dd a) { return dd{-a.hi, -a.lo}; }
NQ_HD dd dd_mul(dd a, dd b) {
  dd p = two_prod(a.hi, b.hi);
  double e = p.lo + (a.hi * b.lo + a.lo * b.hi);
  return fast_two_sum(p.hi, e);
}
NQ_HD dd dd_mul_d(dd a, double b) {
  dd p = two_prod(a.hi, b);
  return fast_two_sum(p.hi, p.lo + a.lo * b);
}
NQ_HD dd dd_div(dd n, dd d) {
  double q1 = n.hi / d.hi;
  dd r = dd_add(n, dd_neg(dd_mul_d(d, q1)));
  double q2 = r.hi / d.hi;
  r = dd_add(r, dd_neg(dd_mul_d(d, q2)));
  double q3 = r.hi / d.hi;
  dd q = fast_two_sum(q1, q2);
  return dd_add_d(q, q3);
}

// ln(x) as a double-double, x finite > 0. Relative error ~2^-64.
NQ_HD dd log_dd(double x) {
  uint64_t u = d2bits(x);
  int e = (int)(u >> 52) & 0x7ff;
  int adj = 0;
  if (e == 0) {  // subnormal: scale up by 2^54
    x *= 18014398509481984.0;
    u = d2bits(x);
    e = (int)(u >> 52) & 0x7ff;
    adj = -54;
  }
  double m = bits2d((u & 0x000fffffffffffffULL) | 0x3ff0000000000000ULL);  // [1,2)
  int ex = e - 1023 + adj;
  if (m > 1.4142135623730951) { m *= 0.5; ex += 1; }  // m in (0.7071, 1.4142]
  double n = m - 1.0;           // exact
  dd den = two_sum(m, 1.0);     // exact
  dd s = dd_div(dd{n, 0.0}, den);
  dd z = dd_mul(s, s);
  double zh = z.hi;
  // R = z^2 * (1/5 + z/7 + ... + z^12/29), plain double (|R| < 2^-12)
  double P = 1.0 / 29.0;
  P = P * zh + 1.0 / 27.0;
  P = P * zh + 1.0 / 25.0;
  P = P * zh + 1.0 / 23.0;
  P = P * zh + 1.0 / 21.0;
  P = P * zh + 1.0 / 19.0;
  P = P * zh + 1.0 / 17.0;
  P = P * zh + 1.0 / 15.0;
  P = P * zh + 1.0 / 13.0;
  P = P * zh + 1.0 / 11.0;
  P = P * zh + 1.0 / 9.0;
  P = P * zh + 1.0 / 7.0;
  P = P * zh + 1.0 / 5.0;
  double R = (zh * zh) * P;
  dd A = dd_mul(z, dd{THIRD_HI, THIRD_LO});
  dd W = two_sum(1.0, A.hi);
  W = fast_two_sum(W.hi, W.lo + (A.lo + R));
  dd L = dd_mul(s, W);
  L.hi *= 2.0; L.lo *= 2.0;
  if (ex == 0) return L;
  dd E = dd_mul_d(dd{LN2_HI, LN2_LO}, (double)ex);
  return dd_add(E, L);
}

// exp core: given t = hi + lo, returns k and S = 2^(j/32) * exp(r) as dd so that exp(t) = 2^k * S.
// p_out (optional) = expm1(r) as dd, j_out the table index.
NQ_HD dd exp_core(dd t, int* k_out, dd* p_out, int* j_out) {
  double fk = rint_(t.hi * INV_LN2_32);
  int kk = (int)fk;
  double rh = t.hi - fk * LN2_32_HI;  // product exact (33-bit constant), difference exact
  dd r = two_sum(rh, t.lo - fk * LN2_32_LO);
  double x = r.hi;
  double h = (x * x) * 0.5;
  double Q = 1.0 / 40320.0;
  Q = Q * x + 1.0 / 5040.0;
  Q = Q * x + 1.0 / 720.0;
  Q = Q * x + 1.0 / 120.0;
  Q = Q * x + 1.0 / 24.0;
  Q = Q * x + 1.0 / 6.0;
  double q = ((x * x) * x) * Q;
  dd p = two_sum(x, r.lo + (h + q));
  int j = kk & 31;
  int k = (kk - j) / 32;
  dd T = dd{exp2_32_tab(j, 0), exp2_32_tab(j, 1)};
  dd Tp = dd_mul(T, p);
  dd S = dd_add(T, Tp);
  *k_out = k;
  if (p_out) *p_out = p;
  if (j_out) *j_out = j;
  return S;
}

NQ_HD double exp_dd(dd t) {
  if (isnan_(t.hi)) return t.hi;
  if (t.hi > 709.79) return bits2d(0x7ff0000000000000ULL);
  if (t.hi < -745.2) return 0.0;
  int k;
  dd S = exp_core(t, &k, nullptr, nullptr);
  return scalbn_(S.hi, k);
}

NQ_HD double nq_exp(double x) { return exp_dd(dd{x, 0.0}); }

NQ_HD double nq_log(double x) {
  dd L = log_dd(x);
  return L.hi;
}

// x^n for small positive integer n by double-double repeated multiplication (x finite)
NQ_HD double powi_dd(double x, int n) {
  dd r = dd{x, 0.0};
  for (int i = 1; i < n; ++i) r = dd_mul_d(r, x);
  return r.hi;
}

// Math.pow for the argument shapes on this path (finite x; any finite or infinite y).
NQ_HD double nq_pow(double x, double y) {
  if (y == 0.0) return 1.0;
  if (isnan_(x) || isnan_(y)) return x + y;
  if (y == 1.0) return x;
  if (y == 2.0) return x * x;
  const double INF = bits2d(0x7ff0000000000000ULL);
  double ax = fabs_(x);
  if (fabs_(y) == INF) {
    if (ax == 1.0) return bits2d(0x7ff8000000000000ULL);  // Java: NaN
    return ((ax > 1.0) == (y > 0.0)) ? INF : 0.0;
  }
  bool yint = (y == rint_(y)) && fabs_(y) < 4503599627370496.0;
  bool yodd = yint && (((long long)y) & 1LL);
  if (ax == 0.0) {
    if (y > 0.0) return (yodd && (d2bits(x) >> 63)) ? -0.0 : 0.0;
    return (yodd && (d2bits(x) >> 63)) ? -INF : INF;
  }
  if (ax == INF) {
    double r = (y > 0.0) ? INF : 0.0;
    return (x < 0.0 && yodd) ? -r : r;
  }
  if (x < 0.0 && !yint) return bits2d(0x7ff8000000000000ULL);
  double sgn = (x < 0.0 && yodd) ? -1.0 : 1.0;
  if (ax == 1.0) return sgn;
  if (yint && y >= 3.0 && y <= 8.0) {
    double r = powi_dd(ax, (int)y);
    if (r != 0.0 && r != INF) return sgn * r;
  }
  dd L = log_dd(ax);
  dd t = dd_mul_d(L, y);
  return sgn * exp_dd(t);
}

// Math.cbrt for x >= 0 finite (counts)
NQ_HD double nq_cbrt(double x) {
  if (x == 0.0) return x;
  double ax = fabs_(x);
  dd L = log_dd(ax);
  dd t = dd_mul(L, dd{THIRD_HI, THIRD_LO});
  double r = exp_dd(t);
  return x < 0.0 ? -r : r;
}

// Math.tanh
NQ_HD double nq_tanh(double x) {
  if (isnan_(x)) return x;
  double t = fabs_(x);
  double r;
  if (t >= 22.0) r = 1.0;
  else if (t < 0x1p-28) r = t;
  else {
    int k, j; dd p;
    dd S = exp_core(dd{2.0 * t, 0.0}, &k, &p, &j);
    dd em1;
    if (k == 0 && j == 0) em1 = p;                       // expm1(2t) directly, no cancellation
    else {
      double sc = pow2i(k);                              // 0 <= k <= 63
      dd E = dd{S.hi * sc, S.lo * sc};
      dd a = two_sum(E.hi, -1.0);
      em1 = fast_two_sum(a.hi, a.lo + E.lo);
    }
    dd den = dd_add_d(em1, 2.0);
    dd q = dd_div(em1, den);
    r = q.hi;
  }
  return (d2bits(x) >> 63) ? -r : r;
}

// argument reduction x = n*pi/2 + r, |x| < 2^20; returns n mod 4 in *q
NQ_HD dd rem_pio2(double x, int* q) {
  double fn = rint_(x * TWO_OVER_PI);
  int n = (int)fn;
  double r0 = x - fn * PIO2_1;                 // exact
  dd a = two_prod(fn, PIO2_2);                 // exact (33-bit constant)
  dd r = two_sum(r0, -a.hi);
  dd b = two_prod(fn, PIO2_3);
  dd r2 = two_sum(r.hi, -b.hi);
  double tail = ((r.lo + r2.lo) - b.lo) - fn * PIO2_3T;
  dd res = fast_two_sum(r2.hi, tail);
  *q = n & 3;
  return res;
}

// sin(r.hi + r.lo), |r| <= pi/4 (+ slack)
NQ_HD double sin_kernel(dd r) {
  double x = r.hi, z = x * x;
  double P = 1.0 / 51090942171709440000.0;       // 1/21!
  P = P * z - 1.0 / 121645100408832000.0;        // 1/19!
  P = P * z + 1.0 / 355687428096000.0;           // 1/17!
  P = P * z - 1.0 / 1307674368000.0;             // 1/15!
  P = P * z + 1.0 / 6227020800.0;                // 1/13!
  P = P * z - 1.0 / 39916800.0;                  // 1/11!
  P = P * z + 1.0 / 362880.0;                    // 1/9!
  P = P * z - 1.0 / 5040.0;                      // 1/7!
  P = P * z + 1.0 / 120.0;                       // 1/5!
  // x^3 * (-1/6 + z*P) with the -x^3/6 term carried in double-double
  dd x2 = two_prod(x, x);
  dd x3 = dd_mul_d(x2, x);
  dd c3 = dd_mul(x3, dd{-0x1.5555555555555p-3, -0x1.5555555555555p-57});  // -1/6
  double tail = (x3.hi * z) * P;
  // cos(x)*r.lo ~ r.lo*(1 - z/2)
  double lo = r.lo * (1.0 - 0.5 * z);
  dd s = two_sum(x, c3.hi);
  double rest = ((s.lo + c3.lo) + tail) + lo;
  return s.hi + rest;
}

// cos(r.hi + r.lo), |r| <= pi/4 (+ slack)
NQ_HD double cos_kernel(dd r) {
  double x = r.hi, z = x * x;
  double P = 1.0 / 2432902008176640000.0;        // 1/20!
  P = P * z - 1.0 / 6402373705728000.0;          // 1/18!
  P = P * z + 1.0 / 20922789888000.0;            // 1/16!
  P = P * z - 1.0 / 87178291200.0;               // 1/14!
  P = P * z + 1.0 / 479001600.0;                 // 1/12!
  P = P * z - 1.0 / 3628800.0;                   // 1/10!
  P = P * z + 1.0 / 40320.0;                     // 1/8!
  P = P * z - 1.0 / 720.0;                       // 1/6!
  // 1 - x^2/2 + x^4/24 + x^6*P - sin(x)*r.lo
  dd x2 = two_prod(x, x);
  dd h = dd{-0.5 * x2.hi, -0.5 * x2.lo};
  dd x4 = dd_mul(x2, x2);
  dd c4 = dd_mul(x4, dd{0x1.5555555555555p-5, 0x1.5555555555555p-59});    // 1/24
  double tail = ((x4.hi * z) * P) - (x * r.lo);
  dd s = two_sum(1.0, h.hi);
  dd s2 = two_sum(s.hi, c4.hi);
  double rest = (((s.lo + s2.lo) + h.lo) + c4.lo) + tail;
  return s2.hi + rest;
}

NQ_HD double nq_sin(double x) {
  if (isnan_(x) || fabs_(x) == bits2d(0x7ff0000000000000ULL)) return bits2d(0x7ff8000000000000ULL);
  if (fabs_(x) < 0x1p-27) return x;
  int q; dd r;
  if (fabs_(x) <= 0.7853981633974483) { q = 0; r = dd{x, 0.0}; }
  else r = rem_pio2(x, &q);
  switch (q) {
    case 0: return sin_kernel(r);
    case 1: return cos_kernel(r);
    case 2: return -sin_kernel(r);
    default: return -cos_kernel(r);
  }
}
NQ_HD double nq_cos(double x) {
  if (isnan_(x) || fabs_(x) == bits2d(0x7ff0000000000000ULL)) return bits2d(0x7ff8000000000000ULL);
  int q; dd r;
  if (fabs_(x) <= 0.7853981633974483) { q = 0; r = dd{x, 0.0}; }
  else r = rem_pio2(x, &q);
  switch (q) {
    case 0: return cos_kernel(r);
    case 1: return -sin_kernel(r);
    case 2: return -cos_kernel(r);
    default: return sin_kernel(r);
  }
}

// atan(t) as dd for t = num/den given as a dd ratio in [0, 1]
NQ_HD dd atan01_dd(dd t) {
  int i = (int)rint_(t.hi * 16.0);
  if (i < 0) i = 0;
  if (i > 16) i = 16;
  double c = (double)i * 0.0625;
  // u = (t - c) / (1 + t*c)
  dd num = dd_add_d(t, -c);
  dd den = dd_add_d(dd_mul_d(t, c), 1.0);
  dd u = (i == 0) ? t : dd_div(num, den);
  double x = u.hi, z = x * x;
  // atan(u) = u + u^3*(-1/3 + z*S), S = 1/5 - z/7 + z^2/9 - z^3/11 + z^4/13 - z^5/15
  double S = -1.0 / 15.0;
  S = S * z + 1.0 / 13.0;
  S = S * z - 1.0 / 11.0;
  S = S * z + 1.0 / 9.0;
  S = S * z - 1.0 / 7.0;
  S = S * z + 1.0 / 5.0;
  dd u2 = dd_mul(u, u);
  dd u3 = dd_mul(u2, u);
  dd c3 = dd_mul(u3, dd{-THIRD_HI, -THIRD_LO});
  double tail = (u3.hi * z) * S;
  dd a = dd_add(u, c3);
  a = dd_add_d(a, tail);
  if (i == 0) return a;
  return dd_add(dd{atan_16_tab(i, 0), atan_16_tab(i, 1)}, a);
}

// Math.atan2(y, x) for finite inputs
NQ_HD double nq_atan2(double y, double x) {
  if (isnan_(x) || isnan_(y)) return x + y;
  bool sy = (d2bits(y) >> 63) != 0, sx = (d2bits(x) >> 63) != 0;
  double ay = fabs_(y), ax = fabs_(x);
  if (ay == 0.0) {
    double r = sx ? PI_HI : 0.0;
    return sy ? -r : r;
  }
  if (ax == 0.0) return sy ? -PIO2_HI : PIO2_HI;
  dd a;
  if (ay <= ax) {
    dd t = dd_div(dd{ay, 0.0}, dd{ax, 0.0});
    a = atan01_dd(t);                       // [0, pi/4]
  } else {
    dd t = dd_div(dd{ax, 0.0}, dd{ay, 0.0});
    a = dd_add(dd{PIO2_HI, PIO2_LO}, dd_neg(atan01_dd(t)));   // pi/2 - atan(x/y)
  }
  if (sx) a = dd_add(dd{PI_HI, PI_LO}, dd_neg(a));
  double r = a.hi;
  if (r == 0.0) r = ay / ax;  // underflow guard
  return sy ? -r : r;
}

}  // namespace nqm
