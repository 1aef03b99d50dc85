// nq_dither.cuh -- palette lookup + error-diffusion along the generalized Hilbert curve, and the
// blue-noise second pass.
// Reference: GilbertCurve (GC:50-373), BlueNoise.diffuse/dither (BN:180-222), nearestColorIndex /
// closestColorIndex (PQ:269-375, PL:329-474), dither() drivers (PQ:393-407, PL:492-522).
//
// The dither is a strictly serial chain along the curve: every pixel needs the quantization error
// of the previous one, the lookups share a first-seen memo when the key is the reduced 16-bit
// colour (PQ:271-274, PL:332-335) and the LAB lookup consumes a java.util.Random stream in visiting
// order (PL:467). So one warp owns one image and walks the curve; inside a step the warp splits the
// work that has no order: the four channels of the error-queue sum (lane & 3), the palette scan
// (palette entry = lane + 32 t) and the three error-shaping channels.
#pragma once
#include "nq_types.h"
#include "nq_color.h"
#include "nq_hist.cuh"
#include "nq_pnn.cuh"

namespace nq {

#define FULL 0xffffffffu

struct WarpShared {
  uint32_t pal[NQ_MAXK];
  float4 palLab[NQ_MAXK];     // alpha, L, A, B of palette[i] (getLab(c2), PL:352)
  float q[NQ_MAXQ][4];        // error queue: FIFO ring, or the PriorityQueue backing array
  double qy[16];              // yDiff of PriorityQueue entries
  float w[NQ_MAXQ];           // current weights[]
  // PnnLABQuantizer.closestColorIndex cost split per channel: T?[v] = cost of a channel difference of
  // |v| (see closest_lab); Ta only when the image is semi-transparent
  double Tr[256], Tg[256], Tb[256], Ta[256];
};

struct Env {
  WarpShared* sh;
  int plen, margin, thresold, DM, ditherMax;
  bool lab, dither, semi, hasTrans, isNano, sorted, useSal, gHasAlpha;
  uint32_t transColor;
  double PR, PG, PB, PA, ratio, gWeight, exp15;
  float beta;
  unsigned short* memo;
  JRandom rng;
  unsigned long long draws;
  int width;
};

__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(FULL, v, src); }

// (dist, idx) argmin across the warp, ties -> larger idx ("last minimal index wins", PQ:291-307)
__device__ __forceinline__ void warp_min_last(double& d, int& i) {
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    double d2 = __shfl_xor_sync(FULL, d, o);
    int i2 = __shfl_xor_sync(FULL, i, o);
    if (d2 < d || (d2 == d && i2 > i)) { d = d2; i = i2; }
  }
}

// PnnQuantizer.nearestColorIndex (PQ:269-311)
__device__ int nearest_rgb(Env& E, uint32_t c) {
  const unsigned lane = lane_id();
  int offset = 0;
  if (E.isNano) {
    offset = color_index(c, E.semi, E.hasTrans);
    unsigned short got = 0;
    if (lane == 0) got = E.memo[offset];
    got = __shfl_sync(FULL, got, 0);
    if (got != 0xFFFF) return got;
  }
  int k = 0;
  if (c_alpha(c) <= 0xF) c = E.transColor;
  if (E.plen > 2 && E.hasTrans && c_alpha(c) > 0xF) k = 1;
  double pr = E.PR, pg = E.PG, pb = E.PB, pa = E.PA;
  if (E.plen < 3) pr = pg = pb = pa = 1;
  const int ca = c_alpha(c), cr = c_red(c), cg = c_green(c), cb = c_blue(c);
  double best = 1e300;
  int bi = -1;
  for (int i = k + lane; i < E.plen; i += 32) {
    const uint32_t c2 = E.sh->pal[i];
    double da = (double)(c_alpha(c2) - ca), dr = (double)(c_red(c2) - cr), dg = (double)(c_green(c2) - cg), db = (double)(c_blue(c2) - cb);
    double cur = pa * (da * da);
    cur += pr * (dr * dr);
    cur += pg * (dg * dg);
    cur += pb * (db * db);
    if (cur <= best) { best = cur; bi = i; }
  }
  warp_min_last(best, bi);
  if (!(best <= 2147483647.0) || bi < 0) bi = k;   // mindist starts at Integer.MAX_VALUE (PQ:286)
  if (E.isNano) {
    if (lane == 0) E.memo[offset] = (unsigned short)bi;
    __syncwarp();
  }
  return bi;
}

// the two smallest (floor(err), index) pairs across the warp; the sequential top-2 scan of
// PQ:329-357 / PL:418-458 keeps exactly these (its comparisons are against truncated ints)
struct Top2 { int i0, i1, d0, d1; };
__device__ __forceinline__ bool t2_less(int da, int ia, int db, int ib) { return da < db || (da == db && ia < ib); }
__device__ __forceinline__ void t2_insert(Top2& t, int d, int i) {
  if (t2_less(d, i, t.d0, t.i0)) { t.d1 = t.d0; t.i1 = t.i0; t.d0 = d; t.i0 = i; }
  else if (t2_less(d, i, t.d1, t.i1)) { t.d1 = d; t.i1 = i; }
}
__device__ __forceinline__ void t2_reduce(Top2& t) {
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    int d0 = __shfl_xor_sync(FULL, t.d0, o), i0 = __shfl_xor_sync(FULL, t.i0, o);
    int d1 = __shfl_xor_sync(FULL, t.d1, o), i1 = __shfl_xor_sync(FULL, t.i1, o);
    t2_insert(t, d0, i0);
    t2_insert(t, d1, i1);
  }
}
// entries whose err truncates to Integer.MAX_VALUE are never stored (err < closest[] fails)
#define T2_NONE 0x7fffffff

// PnnQuantizer.closestColorIndex (PQ:313-375)
__device__ int closest_rgb(Env& E, uint32_t c, int pos) {
  if (c_alpha(c) <= 0xF) return nearest_rgb(E, c);
  const unsigned lane = lane_id();
  double pr = E.PR, pg = E.PG, pb = E.PB, pa = E.PA;
  if (E.plen < 3) pr = pg = pb = pa = 1;
  const int ca = c_alpha(c), cr = c_red(c), cg = c_green(c), cb = c_blue(c);
  Top2 t = {1 << 30, 1 << 30, T2_NONE, T2_NONE};
  for (int k = lane; k < E.plen; k += 32) {
    const uint32_t c2 = E.sh->pal[k];
    double dr = (double)(c_red(c2) - cr), dg = (double)(c_green(c2) - cg), db = (double)(c_blue(c2) - cb);
    double err = pr * (dr * dr);
    err += pg * (dg * dg);
    err += pb * (db * db);
    if (E.semi) { double da = (double)(c_alpha(c2) - ca); err += pa * (da * da); }
    int d = j2i(err);
    if (d != T2_NONE) t2_insert(t, d, k);
  }
  t2_reduce(t);
  int c0 = t.d0 == T2_NONE ? 0 : t.i0, e0 = t.d0;
  int c1 = t.d1 == T2_NONE ? c0 : t.i1, e1 = t.d1;     // PQ:359-360
  int MAX_ERR = E.plen << 2;
  int idx = (pos + 1) % 2;
  if ((double)e1 * .67 < (double)(e1 - e0)) idx = 0;
  else if (c0 > c1) idx = pos % 2;
  const int ci = idx ? c1 : c0, ei = idx ? e1 : e0;
  if (ei >= MAX_ERR || (E.hasTrans && ci == 0)) return nearest_rgb(E, c);
  return ci;
}

// PnnLABQuantizer.nearestColorIndex (PL:329-404)
__device__ int nearest_lab(Env& E, uint32_t c) {
  const unsigned lane = lane_id();
  int offset = 0;
  if (E.isNano) {
    offset = color_index(c, E.semi, E.hasTrans);
    unsigned short got = 0;
    if (lane == 0) got = E.memo[offset];
    got = __shfl_sync(FULL, got, 0);
    if (got != 0xFFFF) return got;
  }
  int k = 0;
  if (c_alpha(c) <= 0xF) c = E.transColor;
  if (E.plen > 2 && E.hasTrans && c_alpha(c) > 0xF) k = 1;
  const int ca = c_alpha(c), cr = c_red(c), cg = c_green(c), cb = c_blue(c);
  const Lab4 l1 = lab_of(c);
  int bi;
  if (E.plen > 4 && !(E.semi || E.plen < 16) && E.plen <= 32) {
    // CIEDE2000 branch (PL:376-395): R_T can be negative, so acceptance depends on the running
    // mindist; one palette entry per lane, resolved in index order
    const int i = k + lane;
    double g = 1e300, f = 1e300;
    if (i < E.plen) {
      const float4 l2 = E.sh->palLab[i];
      double cur = 0;   // hasSemiTransparency is false on this branch
      float tL = ciede_L(l1.L, l2.y);
      cur += (double)tL * (double)tL;
      g = cur;
      CiedeC cc;
      float tC = ciede_C(l1.A, l1.B, l2.z, l2.w, &cc);
      cur += (double)tC * (double)tC;
      g = cur > g ? cur : g;
      double barC, barh;
      float tH = ciede_H(l1.B, l2.w, cc, &barC, &barh);
      cur += (double)tH * (double)tH;
      g = cur > g ? cur : g;
      cur += (double)ciede_RT(barC, barh, tC, tH);
      g = cur > g ? cur : g;
      f = cur;
    }
    double mind = 2147483647.0;
    bi = k;
    unsigned remaining = FULL;
    for (;;) {
      unsigned m = __ballot_sync(FULL, g <= mind) & remaining;
      if (!m) break;
      int L = __ffs(m) - 1;
      mind = shfl_d(f, L);
      bi = k + L;
      remaining = (L == 31) ? 0u : ~((2u << L) - 1u);
    }
  } else {
    const double exp15 = E.exp15;
    double best = 1e300;
    bi = -1;
    for (int i = k + lane; i < E.plen; i += 32) {
      const uint32_t c2 = E.sh->pal[i];
      double cur = 0;
      if (E.semi) { double da = (double)(c_alpha(c2) - ca); cur = (da * da) / exp15; }
      if (E.plen <= 4) {
        double dr = (double)(c_red(c2) - cr), dg = (double)(c_green(c2) - cg), db = (double)(c_blue(c2) - cb);
        cur = (dr * dr) + (dg * dg) + (db * db);
        if (E.semi) { double da = (double)(c_alpha(c2) - ca); cur += da * da; }
      } else {
        const float4 l2 = E.sh->palLab[i];
        if (E.semi || E.plen < 16) {
          double d = (double)(l2.y - l1.L); cur += d * d;
          d = (double)(l2.z - l1.A); cur += d * d;
          d = (double)(l2.w - l1.B); cur += d * d;
        } else {
          cur += (double)fabsf(l2.y - l1.L);
          double da = (double)(l2.z - l1.A), db = (double)(l2.w - l1.B);
          cur += nqm::sqrt_((da * da) + (db * db));
        }
      }
      if (cur <= best) { best = cur; bi = i; }
    }
    warp_min_last(best, bi);
    if (!(best <= 2147483647.0) || bi < 0) bi = k;
  }
  if (E.isNano) {
    if (lane == 0) E.memo[offset] = (unsigned short)bi;
    __syncwarp();
  }
  return bi;
}

// exact cost of palette entry c2 for colour c, in the reference's operation order (PL:421-446)
__device__ __forceinline__ double closest_lab_err(const Env& E, uint32_t c2, int ca, int cr, int cg, int cb) {
  const int ir = c_red(c2) - cr, ig = c_green(c2) - cg, ib = c_blue(c2) - cb;
  double dr = (double)ir, dg = (double)ig, db = (double)ib;
  double err = E.PR * (1 - E.ratio) * (dr * dr);
  err += E.PG * (1 - E.ratio) * (dg * dg);
  err += E.PB * (1 - E.ratio) * (db * db);
  if (E.semi) { double da = (double)(c_alpha(c2) - ca); err += E.PA * (da * da); }
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    double t0 = (double)(c_coeffs[i][0] * (float)ir);   // float * int -> float (PL:437)
    err += E.ratio * (t0 * t0);
    double t1 = (double)(c_coeffs[i][1] * (float)ig);
    err += E.ratio * (t1 * t1);
    double t2 = (double)(c_coeffs[i][2] * (float)ib);
    err += E.ratio * (t2 * t2);
  }
  return err;
}
// per-channel cost tables for the fast scan: every term of closest_lab_err depends on ONE channel
// difference, and all terms are >= 0, so the reference's 12/13-term sequential double sum S and the
// regrouped sum A = Tr[|dr|] + Tg[|dg|] + Tb[|db|] (+ Ta[|da|]) agree to a few ulp:
// |S - A| <= 2e-14 * A. Only floor(S) enters the top-2 decision (see Top2), so A decides it unless A
// lies within 1e-6 (>> 2e-14 * 2^31) of an integer, in which case the exact S is evaluated.
__device__ void closest_lab_tables(const Env& E) {
  for (int v = lane_id(); v < 256; v += 32) {
    const double dv = (double)v;
    double tr = E.PR * (1 - E.ratio) * (dv * dv), tg = E.PG * (1 - E.ratio) * (dv * dv), tb = E.PB * (1 - E.ratio) * (dv * dv);
    for (int i = 0; i < 3; ++i) {
      double t0 = (double)(c_coeffs[i][0] * (float)v), t1 = (double)(c_coeffs[i][1] * (float)v), t2 = (double)(c_coeffs[i][2] * (float)v);
      tr += E.ratio * (t0 * t0); tg += E.ratio * (t1 * t1); tb += E.ratio * (t2 * t2);
    }
    E.sh->Tr[v] = tr; E.sh->Tg[v] = tg; E.sh->Tb[v] = tb;
    E.sh->Ta[v] = E.semi ? E.PA * (dv * dv) : 0.0;
  }
  __syncwarp();
}

// PnnLABQuantizer.closestColorIndex (PL:406-474)
__device__ int closest_lab(Env& E, uint32_t c, int pos) {
  if (c_alpha(c) <= 0xF) return nearest_lab(E, c);
  const unsigned lane = lane_id();
  const int ca = c_alpha(c), cr = c_red(c), cg = c_green(c), cb = c_blue(c);
  Top2 t = {1 << 30, 1 << 30, T2_NONE, T2_NONE};
  for (int k = lane; k < E.plen; k += 32) {
    const uint32_t c2 = E.sh->pal[k];
    double a = E.sh->Tr[abs(c_red(c2) - cr)] + E.sh->Tg[abs(c_green(c2) - cg)] + E.sh->Tb[abs(c_blue(c2) - cb)];
    if (E.semi) a += E.sh->Ta[abs(c_alpha(c2) - ca)];
    int d = j2i(a);
    const double fr = a - (double)d;
    if (fr < 1e-6 || fr > 1.0 - 1e-6) d = j2i(closest_lab_err(E, c2, ca, cr, cg, cb));
    if (d != T2_NONE) t2_insert(t, d, k);
  }
  t2_reduce(t);
  const int c0 = t.d0 == T2_NONE ? 0 : t.i0, e0 = t.d0;
  const int c1 = t.d1 == T2_NONE ? c0 : t.i1, e1 = t.d1;     // PL:460-461
  int idx = 1;
  if (e0 == 0) idx = 0;                                       // short-circuit: no draw (PL:467)
  else {
    int r = E.rng.next_int(32767);
    ++E.draws;
    int sum = (int)((unsigned)e1 + (unsigned)e0);
    if ((r % sum) <= e1) idx = 0;
  }
  const int ci = idx ? c1 : c0, ei = idx ? e1 : e0;
  int MAX_ERR = E.plen;
  if (ei >= MAX_ERR || ci == 0 || c_alpha(E.sh->pal[ci]) < ca) return nearest_lab(E, c);
  return ci;
}

// Ditherable.nearestColorIndex as bound by getDitherFn (PQ:377-391, PL:476-490)
__device__ __forceinline__ int lookup(Env& E, uint32_t c, int pos) {
  if (E.lab) return E.plen <= 4 ? nearest_lab(E, c) : closest_lab(E, c, pos);
  return E.dither ? nearest_rgb(E, c) : closest_rgb(E, c, pos);
}

// GilbertCurve.normalDistribution (GC:114-123)
__device__ float normal_distribution(float x, float peak) {
  const float mean = .5f, stdDev = .1f;
  double d = (double)(x - mean);
  double sd = (double)stdDev;
  double exponent = -(d * d) / (2 * (sd * sd));
  double pdf = (1 / (sd * nqm::sqrt_(2 * 3.141592653589793))) * nqm::nq_exp(exponent);
  double maxPdf = 1 / (sd * nqm::sqrt_(2 * 3.141592653589793));
  double scaledPdf = (pdf / maxPdf) * (double)peak;
  return (float)dmax(0.0, dmin((double)peak, scaledPdf));
}

// Y_Diff(pixel, c) with color2Y(pixel) already known (CL:215-227)
__device__ __forceinline__ double y_diff_pre(double ypix, uint32_t c, const double* lut) {
  return nqm::fabs_(color_y(c, lut) - ypix) * 100;
}

// GilbertCurve.ditherPixel (GC:125-185). qcur = qPixels[bidx] at the time of the call; sal =
// saliencies[bidx]; ypix = color2Y(pixel).
__device__ int dither_pixel(Env& E, int x, int y, int bidx, uint32_t pixel, float sal, double ypix, uint32_t c2, float beta, int qcur) {
  const int plen = E.plen, margin = E.margin;
  const double weight = E.gWeight;
  const uint32_t qcol = E.sh->pal[qcur];
  const signed char* bn = g_blueNoise;
  const double* lut = g_gammaLut;
  const int r_pix = c_red(c2), g_pix = c_green(c2), b_pix = c_blue(c2), a_pix = c_alpha(c2);
  const float strength = 1 / 3.f;
  const int acceptedDiff = max(2, plen - margin);
  if (plen <= 4 && sal > .2f && sal < .25f)
    c2 = bn_diffuse(pixel, qcol, beta * 2 / sal, strength, x, y, bn);
  else if (plen <= 4 || y_diff_pre(ypix, c2, lut) < (double)(2 * acceptedDiff)) {
    if (plen > 64) {
      float kappa = sal < .6f ? beta * .15f / sal : beta * .4f / sal;
      c2 = bn_diffuse(pixel, qcol, kappa, strength, x, y, bn);
    } else if (plen > 16 && weight < .005)
      c2 = bn_diffuse(pixel, qcol, beta * normal_distribution(sal, .5f) + beta, strength, x, y, bn);
    else
      c2 = bn_diffuse(pixel, qcol, beta * .5f / sal, strength, x, y, bn);
  }

  double gamma = (plen <= 32 && weight < .01 && weight > .007) ? (double)(1 - beta) : (double)beta;
  if (plen > 4 && y_diff_pre(ypix, c2, lut) > (gamma * acceptedDiff)) {
    if (margin > 6 || gamma > (double)beta) {
      float kappa = sal < .4f ? beta * .4f * sal : beta * .4f / sal;
      uint32_t c1 = c_argb(a_pix, r_pix, g_pix, b_pix);
      if (plen > 32 && (double)sal < .9)
        kappa = beta * normal_distribution(sal, 2.f);
      else {
        if (weight >= .0015 && (double)sal < .6) c1 = pixel;
        if (weight >= .005 && (double)sal < .6)
          kappa = beta * normal_distribution(sal, weight < .0008 ? 2.5f : 1.75f);
        else if (plen >= 32 || y_diff(c1, c2, lut) > (gamma * 3.141592653589793 * acceptedDiff)) {
          double ub = 1 - plen / 320.0;
          if ((double)sal > .15 && (double)sal < ub)
            kappa = beta * (!E.sorted && weight < .0025 ? .55f : .5f) / sal;
          else
            kappa = beta * normal_distribution(sal, weight < .0025 ? 1.82f : 2.f);
        }
      }
      c2 = bn_diffuse(c1, qcol, kappa, strength, x, y, bn);
    } else if (plen <= 32 && weight >= .004)
      c2 = bn_diffuse(c2, qcol, beta * normal_distribution(sal, .25f), strength, x, y, bn);
    else
      c2 = c_argb(a_pix, r_pix, g_pix, b_pix);
  }

  if (E.DM < 16 && plen > 4 && sal < .6f && y_diff_pre(ypix, c2, lut) > (double)(margin - 1))
    c2 = c_argb(a_pix, r_pix, g_pix, b_pix);
  if (plen > 32 && (double)sal > .95) {
    float kappa = beta * fmaxf(.05f, .75f - plen / 128.f) * sal;
    c2 = bn_diffuse(pixel, qcol, kappa, strength, x, y, bn);
  }
  return lookup(E, c2, bidx);
}

// (float) Math.tanh(x) for the error shaping (GC:255). The double result is only used after narrowing
// to float, so a plain-double evaluation t with relative error < 1e-13 decides the float whenever
// (float)(t(1-d)) == (float)(t(1+d)), d = 1e-12 (rounding is monotone); otherwise, and for |x| < 1
// where 1 - 2/(e^2x + 1) cancels, the double-double kernel runs. Same float as (float)nq_tanh(x).
__device__ __forceinline__ float tanh_to_float(double x) {
  const double ax = nqm::fabs_(x);
  if (ax >= 1.0 && ax < 22.0) {
    const double y = 2.0 * ax;
    const double fk = nqm::rint_(y * nqm::INV_LN2_32);
    const int kk = (int)fk;
    const double r = (y - fk * nqm::LN2_32_HI) - fk * nqm::LN2_32_LO;
    double p = 1.0 / 720.0;
    p = p * r + 1.0 / 120.0;
    p = p * r + 1.0 / 24.0;
    p = p * r + 1.0 / 6.0;
    p = p * r + 0.5;
    p = p * r + 1.0;
    p = p * r;                                     // expm1(r), |r| <= 0.011
    const double T = nqm::exp2_32_tab(kk & 31, 0);
    const double Ee = (T + T * p) * nqm::pow2i(kk >> 5);
    const double t = 1.0 - 2.0 / (Ee + 1.0);
    const float lo = (float)(t * (1.0 - 1e-12)), hi = (float)(t * (1.0 + 1e-12));
    if (lo == hi) return x < 0 ? -lo : lo;
  }
  return (float)nqm::nq_tanh(x);
}

// java.util.PriorityQueue over the shared arrays, comparator Double.compare(o2.yDiff, o1.yDiff)
// (GC:87-94): "x before e" <=> x.yDiff > e.yDiff. All lanes run the same steps; lane 0 stores.
struct PQ {
  WarpShared* sh;
  int n;
  __device__ __forceinline__ void put(int k, const float* p, double yd) {
    if (lane_id() == 0) { sh->q[k][0] = p[0]; sh->q[k][1] = p[1]; sh->q[k][2] = p[2]; sh->q[k][3] = p[3]; sh->qy[k] = yd; }
  }
  __device__ __forceinline__ void move(int dst, int src) {
    float p[4] = {sh->q[src][0], sh->q[src][1], sh->q[src][2], sh->q[src][3]};
    double yd = sh->qy[src];
    __syncwarp();
    put(dst, p, yd);
    __syncwarp();
  }
  __device__ void offer(const float* p, double yd) {   // siftUp
    int k = n++;
    while (k > 0) {
      int parent = (k - 1) >> 1;
      double pe = sh->qy[parent];
      if (!(yd > pe)) break;            // cmp(x, e) >= 0
      move(k, parent);
      k = parent;
    }
    __syncwarp();
    put(k, p, yd);
    __syncwarp();
  }
  __device__ void poll() {              // remove head, siftDown the last element
    int last = --n;
    if (last == 0) return;
    float p[4] = {sh->q[last][0], sh->q[last][1], sh->q[last][2], sh->q[last][3]};
    double yd = sh->qy[last];
    __syncwarp();
    int k = 0, half = last >> 1;
    while (k < half) {
      int child = (k << 1) + 1, right = child + 1;
      double cy = sh->qy[child];
      if (right < last) { double ry = sh->qy[right]; if (ry > cy) { child = right; cy = ry; } }   // cmp(c, right) > 0
      if (!(cy > yd)) break;            // cmp(x, c) <= 0
      move(k, child);
      k = child;
    }
    put(k, p, yd);
    __syncwarp();
  }
};

// -------------------------------------------------------------------------------------------------
// GilbertCurve constructor constants + initWeights (GC:50-112, 336-354), one thread per image
// -------------------------------------------------------------------------------------------------
__device__ void init_weights(float* weights, int size) {
  const float weightRatio = (float)nqm::nq_pow((double)(343.f + 1.f), (double)(1.f / ((float)size - 1.f)));
  float weight = 1.f, sumweight = 0.f;
  for (int c = 0; c < size; ++c) {
    sumweight += (weights[size - c - 1] = weight);
    weight /= weightRatio;
  }
  weight = 0.f;
  for (int c = 0; c < size; ++c) weight += (weights[c] /= sumweight);
  weights[0] += 1.f - weight;
}

__global__ void k_dither_setup(NqImage* imgs, const NqSlot* slots, int nimg) {
  int img = blockIdx.x * blockDim.x + threadIdx.x;
  if (img >= nimg) return;
  NqImage& I = imgs[img];
  if (I.nmax <= 2) {   // fixed palette (PQ:441-452)
    I.weight = 1;
    I.paletteLen = I.nmax;
    if (I.transIdx >= 0) { I.palette[0] = I.transColor; I.palette[1] = 0xFF000000u; }
    else { I.palette[0] = 0xFF000000u; I.palette[1] = 0xFFFFFFFFu; }
    if (I.nmax < 2) I.paletteLen = I.nmax < 0 ? 0 : I.nmax;
  }
  const int plen = I.paletteLen;
  const bool lab = I.kind == NQ_KIND_LAB;
  const double weight = I.hasSemi ? -I.weight : I.weight;   // PQ:396-397, PL:496-497
  // PnnQuantizer.nearestColorIndex tests the FIELD `weight > .015` while it carries the sign flip
  // (PQ:271 after PQ:396-397): a semi-transparent image always uses the reduced memo key there.
  // PnnLABQuantizer latched isNano inside pnnquan, before the flip (PL:180).
  if (!lab) I.isNano = !(weight > .015);
  // which saliency map exists (PL:135, PL:499-508)
  const bool salFromPnn = lab && I.nmax > 2 && I.nmax < 128;
  const bool salFromDither = lab && I.dither && !salFromPnn && (plen <= 256 || weight > .99);
  const bool sal = salFromPnn || salFromDither;
  I.gUseSal = sal;
  const bool hasAlpha = weight < 0;
  I.gHasAlpha = hasAlpha;
  I.gWeight = nqm::fabs_(weight);
  I.gMargin = weight < .0025 ? 12 : weight < .004 ? 8 : 6;
  const bool sorted = plen > 128 && weight >= .02 && (!hasAlpha || weight < .18);
  I.gSorted = sorted;
  float beta = plen > 4 ? (float)(.6f - .00625f * (float)plen) : 1.f;
  if (plen > 4) {
    double boundary = .005 - .0000625 * plen;
    beta = (float)(weight > boundary ? .25 : dmin(1.5, (double)beta + plen * weight));
    if (plen > 16 && plen <= 32 && weight < .003) beta += .075f;
    else if (weight < .0015 || (plen > 32 && plen < 256)) beta += .1f;
    if ((plen >= 64 && (weight > .012 && weight < .0125)) || (weight > .025 && weight < .03)) beta += .05f;
    else if (plen > 32 && plen < 64 && weight < .015) beta = .55f;
    else if (plen > 16 && plen <= 32 && weight <= .005) beta += (float)(.05 + weight * plen);
  } else
    beta *= .95f;
  if (plen > 64 || (plen > 4 && weight > .02)) beta *= .4f;
  if (plen > 64 && weight < .02) beta = .18f;
  int DM = weight < .015 ? ((weight > .0025) ? 25 : 16) : 9;
  if (weight > .99) { beta = (float)weight; DM = 25; }
  double edge = hasAlpha ? 1 : nqm::nq_exp(weight) - .25;
  double deviation = weight > .002 ? -.25 : 1;
  double sq = nqm::sqrt_((double)DM) + edge * deviation;
  int ditherMax = (hasAlpha || DM > 9) ? j2b(sq * sq) : j2b(DM * (sal ? 2.0 : 2.718281828459045));
  const int density = plen > 16 ? 3200 : 1500;
  if (plen / weight > 5000 && (weight > .045 || (weight > .01 && plen < 64))) ditherMax = j2b((5 + edge) * (5 + edge));
  else if (weight < .03 && plen / weight < density && plen >= 16 && plen < 256) ditherMax = j2b((5 + edge) * (5 + edge));
  I.gBeta = beta; I.gDitherMaxQ = DM; I.gDitherMax = ditherMax;
  I.gThresold = DM > 9 ? -112 : -64;
  init_weights(I.gWeights, DM);
  init_weights(I.gW1, 1);
  init_weights(I.gW3, 3);
  init_weights(I.gW7, 7);
  I.bnWeight = 1.0f;
}

// -------------------------------------------------------------------------------------------------
// the serial pass: Gilbert-order error diffusion (GC:187-280), then BlueNoise.dither (BN:207-222)
// when dither == false and the palette has more than 32 entries. One warp per image.
//
// The warp walks the curve in blocks of 32 pixels. Everything that does not depend on the running
// error is gathered for a whole block at once, one pixel per lane, a block ahead of its use: the
// visiting order, the source pixel, its saliency (from the Lab table) and its luminance. Results are
// kept one per lane and written back per block. Only the serial recurrence stays on the chain.
// -------------------------------------------------------------------------------------------------
struct PixBlock { uint32_t xy, px; float sal; double ypix; };

__device__ __forceinline__ int chan(uint32_t c, int ch) {   // ErrorBox channel order r, g, b, a (GC:24-31)
  return (int)((c >> (ch == 3 ? 24 : 16 - 8 * ch)) & 0xFF);
}

__global__ void __launch_bounds__(32) k_dither(NqImage* imgs, const NqSlot* slots, const uint32_t* order) {
  __shared__ WarpShared sh;
  const int img = blockIdx.x;
  NqImage& I = imgs[img];
  const NqSlot& S = slots[img];
  const unsigned lane = lane_id();
  const int npix = I.npix, width = I.width, plen = I.paletteLen;
  if (plen <= 0 || I.error) return;
  const uint32_t* in = S.in;
  uint32_t* out = S.out;
  const int fixA0 = I.fixA0;

  Env E;
  E.sh = &sh;
  E.plen = plen; E.margin = I.gMargin; E.thresold = I.gThresold; E.DM = I.gDitherMaxQ; E.ditherMax = I.gDitherMax;
  E.lab = I.kind == NQ_KIND_LAB; E.dither = I.dither != 0; E.semi = I.hasSemi != 0; E.hasTrans = I.transIdx >= 0;
  E.isNano = I.isNano != 0; E.sorted = I.gSorted != 0; E.useSal = I.gUseSal != 0; E.gHasAlpha = I.gHasAlpha != 0;
  E.transColor = I.transColor;
  E.PR = I.PR; E.PG = I.PG; E.PB = I.PB; E.PA = I.PA; E.ratio = I.ratioMerge; E.gWeight = I.gWeight;
  E.exp15 = E.semi ? nqm::nq_exp(1.5) : 1.0;
  E.beta = I.gBeta;
  E.memo = S.memo;
  E.rng.set_seed(I.seed);
  E.draws = 0;
  E.width = width;

  for (int i = lane; i < plen; i += 32) {
    uint32_t pc = I.palette[i];
    sh.pal[i] = pc;
    if (E.lab) { Lab4 l = lab_of(pc); sh.palLab[i] = make_float4(l.alpha, l.L, l.A, l.B); }
  }
  const int DM = E.DM;
  if (lane < NQ_MAXQ) sh.w[lane] = lane < (unsigned)DM ? I.gWeights[lane] : 0.f;
  for (int i = lane; i < NQ_MAXQ * 4; i += 32) sh.q[i >> 2][i & 3] = 0.f;
  if (lane < 16) sh.qy[lane] = 0;
  __syncwarp();
  if (E.lab && plen > 4) closest_lab_tables(E);

  const bool sorted = E.sorted, useSal = E.useSal, dither = E.dither, gHasAlpha = E.gHasAlpha;
  const int ch = lane & 3;
  const int thresold = E.thresold, ditherMax = E.ditherMax, margin = E.margin;
  const float beta = E.beta;
  const double* lut = g_gammaLut;
  const signed char* bn = g_blueNoise;
  const bool salReplaced = I.nmax < 128 && I.nmax > 2;   // which pixel feeds getLab for the saliency (PL:141-156 vs PL:503-506)
  const uint32_t transColor = I.transColor;
  int head = 0;            // FIFO: slot of the oldest entry (the queue always holds DM entries)
  PQ pq{&sh, 0};
  int wlen = 0;            // weights.length in sorted mode (0, 1, 3, 7)

  auto fetch = [&](int n0) {
    PixBlock b;
    b.xy = 0; b.px = 0; b.sal = 0.f; b.ypix = 0.0;
    const int n = n0 + (int)lane;
    if (n < npix) {
      b.xy = order[n];
      const int bidx = (int)(b.xy & 0xFFFF) + (int)(b.xy >> 16) * width;
      b.px = eff_pixel(in[bidx], fixA0);
      if (useSal) {
        uint32_t sp = b.px;
        if (salReplaced && (sp >> 24) <= 0xF) sp = transColor;
        const Lab4 l = lab_of(sp);
        const float saliencyBase = .1f;
        b.sal = saliencyBase + (1 - saliencyBase) * l.L / 100.f * l.alpha / 255.f;
      }
      b.ypix = color_y(b.px, lut);
    }
    return b;
  };

  PixBlock nxt = fetch(0);
  for (int n0 = 0; n0 < npix; n0 += 32) {
    const PixBlock cur = nxt;
    if (n0 + 32 < npix) nxt = fetch(n0 + 32);
    uint32_t myOut = 0;
    const int cnt = min(32, npix - n0);
    for (int j = 0; j < cnt; ++j) {
      const uint32_t xy = __shfl_sync(FULL, cur.xy, j);
      const uint32_t pixel = __shfl_sync(FULL, cur.px, j);
      const float sal = __shfl_sync(FULL, cur.sal, j);
      const double ypix = shfl_d(cur.ypix, j);
      const int x = xy & 0xFFFF, y = xy >> 16, bidx = x + y * width;

      // ---- error.p = pixel + sum(queue[i].p * weights[i]) in queue order (GC:190-204); lane&3 = channel
      float acc = (float)chan(pixel, ch);
      float mx = (float)(DM - 1);
      if (!sorted) {
        int slot = head;
        for (int i = 0; i < DM; ++i) {
          acc += sh.q[slot][ch] * sh.w[i];
          if (acc > mx) mx = acc;
          if (++slot == DM) slot = 0;
        }
      } else {
        int i = wlen - 1;
        for (int qi = 0; qi < pq.n && i >= 0; ++qi, --i) {
          acc += sh.q[qi][ch] * sh.w[i];
          if (acc > mx) mx = acc;
        }
      }
      mx = fmaxf(mx, __shfl_xor_sync(FULL, mx, 1));
      mx = fmaxf(mx, __shfl_xor_sync(FULL, mx, 2));
      const float maxErr = mx;
      const int mine = j2i(dmin(255.0, dmax((double)acc, 0.0)));
      const int r_pix = __shfl_sync(FULL, mine, 0), g_pix = __shfl_sync(FULL, mine, 1), b_pix = __shfl_sync(FULL, mine, 2), a_pix = __shfl_sync(FULL, mine, 3);

      // ---- quantize (GC:211-229)
      uint32_t c2 = c_argb(a_pix, r_pix, g_pix, b_pix);
      int qi;
      if (useSal && dither && !sorted && (!gHasAlpha || c_alpha(pixel) < a_pix)) {
        if ((plen >= 256 && sal > .99f) || (gHasAlpha && (double)(c_alpha(pixel) - a_pix) < (.5 * margin)))
          qi = lookup(E, c2, bidx);
        else
          qi = dither_pixel(E, x, y, bidx, pixel, sal, ypix, c2, beta, 0);    // qPixels[bidx] is still 0 here (GC:136,216)
      } else if (plen <= 32 && a_pix > 0xF0) {
        qi = lookup(E, c2, bidx);
        const int acceptedDiff = max(2, plen - margin);
        if (useSal && (y_diff_pre(ypix, c2, lut) > (double)acceptedDiff || u_diff(pixel, c2) > (double)(2 * acceptedDiff))) {
          const float strength = 1 / 3.f;
          c2 = bn_diffuse(pixel, sh.pal[qi], 1 / sal, strength, x, y, bn);
          qi = lookup(E, c2, bidx);
        }
      } else
        qi = lookup(E, c2, bidx);

      // ---- queue maintenance (GC:231-234). FIFO: the size is always DITHER_MAX, poll() drops the
      //      oldest entry and its slot receives the new error below.
      if (sorted) {
        if (pq.n >= DM) pq.poll();
        else if (pq.n != 0) {
          const int size = pq.n;                      // initWeights(size): size empty boxes + new weights
          const float zero[4] = {0.f, 0.f, 0.f, 0.f};
          for (int k = 0; k < size; ++k) pq.offer(zero, 0.0);
          const float* src = size == 1 ? I.gW1 : (size == 3 ? I.gW3 : I.gW7);
          if (lane < (unsigned)size) sh.w[lane] = src[lane];
          wlen = size;
          __syncwarp();
        }
      }

      // ---- error of this pixel and its shaping (GC:236-264); lanes 0..2 shape r, g, b
      c2 = sh.pal[qi];
      const int pixch = ch == 0 ? r_pix : (ch == 1 ? g_pix : (ch == 2 ? b_pix : a_pix));
      float e = (float)(pixch - chan(c2, ch));
      const bool denoise = plen > 2;
      const bool diffuse = bn[bidx & 4095] > thresold;
      const double yDiff = sorted ? y_diff_pre(ypix, c2, lut) : 1.0;
      const bool illusion = !diffuse && bn[j2i(yDiff * 4096) & 4095] > thresold;
      bool unacc = false;
      if (denoise && ch < 3) {
        if (fabsf(e) >= (float)ditherMax) {
          if (sorted && useSal) unacc = true;
          if (diffuse) e = tanh_to_float((double)(e / maxErr * 20.f)) * (float)(ditherMax - 1);
          else if (illusion) e = (float)((double)(e / maxErr) * yDiff) * (float)(ditherMax - 1);
          else e /= (float)(1 + nqm::sqrt_((double)ditherMax));
        }
        if (sorted && !useSal && fabsf(e) >= (float)DM) unacc = true;
      }
      const bool unaccepted = (__ballot_sync(FULL, unacc) & 7u) != 0;

      if (unaccepted) {   // GC:266-274
        if (useSal) qi = dither_pixel(E, x, y, bidx, pixel, sal, ypix, c2, beta, qi);
        else if (y_diff_pre(ypix, c2, lut) > 3 && u_diff(pixel, c2) > 3) {
          const float strength = 1 / 3.f;
          c2 = bn_diffuse(pixel, sh.pal[qi], strength, strength, x, y, bn);
          qi = lookup(E, c2, bidx);
        }
      }

      // ---- errorq.add(error) (GC:276)
      if (!sorted) {
        __syncwarp();
        if (lane < 4) sh.q[head][ch] = e;
        if (++head == DM) head = 0;
        __syncwarp();
      } else {
        float p[4];
        p[0] = __shfl_sync(FULL, e, 0); p[1] = __shfl_sync(FULL, e, 1); p[2] = __shfl_sync(FULL, e, 2); p[3] = __shfl_sync(FULL, e, 3);
        pq.offer(p, yDiff);
      }

      const uint32_t res = (dither || plen <= 32) ? sh.pal[qi] : (uint32_t)qi;   // GC:278-279
      if ((int)lane == j) myOut = res;
    }
    if ((int)lane < cnt) {
      const int bidx = (int)(cur.xy & 0xFFFF) + (int)(cur.xy >> 16) * width;
      out[bidx] = myOut;
    }
  }

  // ---- BlueNoise.dither second pass (PQ:400-401, PL:511-515, BN:207-222); memo and RNG carry over
  if (!dither && plen > 32) {
    __syncwarp();
    __threadfence_block();
    const float weight = I.bnWeight, strength = 1 / 3.f;
    for (int n0 = 0; n0 < npix; n0 += 32) {
      const int n = n0 + (int)lane;
      uint32_t px = 0, qv = 0;
      if (n < npix) { px = eff_pixel(in[n], fixA0); qv = out[n]; }
      uint32_t myOut = 0;
      const int cnt = min(32, npix - n0);
      for (int j = 0; j < cnt; ++j) {
        const int bidx = n0 + j, x = bidx % width, y = bidx / width;
        const uint32_t pixel = __shfl_sync(FULL, px, j);
        const uint32_t q0 = __shfl_sync(FULL, qv, j);
        const uint32_t c1 = bn_diffuse(pixel, sh.pal[q0], weight, strength, x, y, bn);
        const int qi = lookup(E, c1, bidx);
        if ((int)lane == j) myOut = sh.pal[qi];
      }
      if (n < npix) out[n] = myOut;
    }
  }
  if (lane == 0) I.rngDraws = E.draws;
}

}  // namespace nq
