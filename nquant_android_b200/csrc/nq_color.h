// nq_color.h -- colour arithmetic of the quantizer path, written once for device and host.
// Each routine cites the reference lines whose arithmetic (operand types, evaluation order, casts)
// it reproduces. All transcendental calls go through nq_math.h.
#pragma once
#include "nq_math.h"

namespace nq {

struct Lab4 { float alpha, L, A, B; };

// ---- android.graphics.Color ------------------------------------------------------------------
NQ_HD int c_alpha(uint32_t c) { return (int)(c >> 24); }
NQ_HD int c_red(uint32_t c) { return (int)((c >> 16) & 0xFF); }
NQ_HD int c_green(uint32_t c) { return (int)((c >> 8) & 0xFF); }
NQ_HD int c_blue(uint32_t c) { return (int)(c & 0xFF); }
NQ_HD uint32_t c_argb(int a, int r, int g, int b) {
  return ((uint32_t)a << 24) | ((uint32_t)r << 16) | ((uint32_t)g << 8) | (uint32_t)b;
}

// Java (int) cast of a double (JLS 5.1.3): truncation, saturation, NaN -> 0
NQ_HD int j2i(double v) {
#if defined(__CUDA_ARCH__)
  return __double2int_rz(v);
#else
  if (v != v) return 0;
  if (v >= 2147483647.0) return 2147483647;
  if (v <= -2147483648.0) return (int)0x80000000;
  return (int)v;
#endif
}
NQ_HD int j2b(double v) { return (int)(signed char)(unsigned char)(unsigned)j2i(v); }
NQ_HD double dmin(double a, double b) { return a < b ? a : b; }
NQ_HD double dmax(double a, double b) { return a > b ? a : b; }
NQ_HD float fabsf_(float x) { return x < 0.f ? -x : (x == 0.f ? 0.f : x); }
// Math.round(double) for |v| < 2^62
NQ_HD long long jround(double v) {
  double f = floor(v);
  if (v - f >= 0.5) f += 1.0;
  return (long long)f;
}

// BitmapUtilities.getColorIndex (BU:8-15)
NQ_HD int color_index(uint32_t c, bool semi, bool transp) {
  if (semi) return (c_alpha(c) & 0xF0) << 8 | (c_red(c) & 0xF0) << 4 | (c_green(c) & 0xF0) | (c_blue(c) >> 4);
  if (transp) return (c_alpha(c) & 0x80) << 8 | (c_red(c) & 0xF8) << 7 | (c_green(c) & 0xF8) << 2 | (c_blue(c) >> 3);
  return (c_red(c) & 0xF8) << 8 | (c_green(c) & 0xFC) << 3 | (c_blue(c) >> 3);
}

// ---- sRGB linearisation table: lut[v] = gammaToLinear(v) (CL:71-75; same formula inside
//      androidx ColorUtils.RGBToXYZ) ---------------------------------------------------------------
NQ_HD double gamma_to_linear(int channel) {
  const double c = channel / 255.0;
  return c < 0.04045 ? c / 12.92 : nqm::nq_pow((c + 0.055) / 1.055, 2.4);
}

// androidx ColorUtils.pivotXyzComponent
NQ_HD double pivot_xyz(double component) {
  return component > 0.008856 ? nqm::nq_pow(component, 1 / 3.0) : (903.3 * component + 16) / 116;
}

// CIELABConvertor.RGB2LAB (CL:58-69) over androidx ColorUtils.colorToLAB
NQ_HD Lab4 rgb2lab(uint32_t c, const double* lut) {
  double sr = lut[c_red(c)], sg = lut[c_green(c)], sb = lut[c_blue(c)];
  double x = 100 * (sr * 0.4124 + sg * 0.3576 + sb * 0.1805);
  double y = 100 * (sr * 0.2126 + sg * 0.7152 + sb * 0.0722);
  double z = 100 * (sr * 0.0193 + sg * 0.1192 + sb * 0.9505);
  x = pivot_xyz(x / 95.047);
  y = pivot_xyz(y / 100.0);
  z = pivot_xyz(z / 108.883);
  Lab4 o;
  o.alpha = (float)c_alpha(c);
  o.L = (float)dmax(0.0, 116 * y - 16);
  o.A = (float)(500 * (x - y));
  o.B = (float)(200 * (y - z));
  return o;
}

NQ_HD int constrain255(long long v) { return v < 0 ? 0 : (v > 255 ? 255 : (int)v); }

// CIELABConvertor.LAB2RGB (CL:77-80) over ColorUtils.LABToColor + setAlphaComponent.
// alpha is the already truncated (int) value; returns false when setAlphaComponent would throw.
NQ_HD bool lab2rgb(int alpha, float Lf, float Af, float Bf, uint32_t* out) {
  const double l = Lf, a = Af, b = Bf;
  const double fy = (l + 16) / 116;
  const double fx = a / 500 + fy;
  const double fz = fy - b / 200;
  double tmp = nqm::nq_pow(fx, 3);
  const double xr = tmp > 0.008856 ? tmp : (116 * fx - 16) / 903.3;
  const double yr = l > 903.3 * 0.008856 ? nqm::nq_pow(fy, 3) : l / 903.3;
  tmp = nqm::nq_pow(fz, 3);
  const double zr = tmp > 0.008856 ? tmp : (116 * fz - 16) / 903.3;
  const double x = xr * 95.047, y = yr * 100.0, z = zr * 108.883;
  double r = (x * 3.2406 + y * -1.5372 + z * -0.4986) / 100;
  double g = (x * -0.9689 + y * 1.8758 + z * 0.0415) / 100;
  double bb = (x * 0.0557 + y * -0.2040 + z * 1.0570) / 100;
  r = r > 0.0031308 ? 1.055 * nqm::nq_pow(r, 1 / 2.4) - 0.055 : 12.92 * r;
  g = g > 0.0031308 ? 1.055 * nqm::nq_pow(g, 1 / 2.4) - 0.055 : 12.92 * g;
  bb = bb > 0.0031308 ? 1.055 * nqm::nq_pow(bb, 1 / 2.4) - 0.055 : 12.92 * bb;
  int ri = constrain255(jround(r * 255)), gi = constrain255(jround(g * 255)), bi = constrain255(jround(bb * 255));
  if (alpha < 0 || alpha > 255) return false;
  *out = c_argb(alpha, ri, gi, bi);
  return true;
}

// ---- CIEDE2000 pieces (CL:86-194) ---------------------------------------------------------------
NQ_HD float deg2rad(double deg) { return (float)(deg * (3.141592653589793 / 180.0)); }

// L' term (CL:91-98)
NQ_HD float ciede_L(float L1, float L2) {
  float deltaLPrime = L2 - L1;
  float barLPrime = (L1 + L2) / 2.f;
  double d = (double)(barLPrime - 50.f);
  double d2 = d * d;  // Math.pow(x, 2)
  float S_L = (float)(1 + (((double)0.015f * d2) / nqm::sqrt_(20 + d2)));
  return deltaLPrime / (1.0f * S_L);
}

struct CiedeC { double a1p, a2p, C1p, C2p; };
// C' term (CL:100-118)
NQ_HD float ciede_C(float A1, float B1, float A2, float B2, CiedeC* o) {
  const float pow25To7 = 6103515625.f;
  float C1 = (float)nqm::sqrt_((double)((A1 * A1) + (B1 * B1)));
  float C2 = (float)nqm::sqrt_((double)((A2 * A2) + (B2 * B2)));
  float barC = (C1 + C2) / 2.f;
  double p7 = nqm::nq_pow((double)barC, 7.0);
  float G = (float)((double)0.5f * (1 - nqm::sqrt_(p7 / (p7 + (double)pow25To7))));
  o->a1p = (1.0 + (double)G) * (double)A1;
  o->a2p = (1.0 + (double)G) * (double)A2;
  o->C1p = nqm::sqrt_((o->a1p * o->a1p) + (double)(B1 * B1));
  o->C2p = nqm::sqrt_((o->a2p * o->a2p) + (double)(B2 * B2));
  float deltaCPrime = (float)o->C2p - (float)o->C1p;
  float barCPrime = ((float)o->C1p + (float)o->C2p) / 2.f;
  float S_C = 1 + (0.045f * barCPrime);
  return deltaCPrime / (1.f * S_C);
}

// H' term (CL:120-185)
NQ_HD float ciede_H(float B1, float B2, const CiedeC& c, double* barCPrime, double* barhPrime) {
  const float deg360 = deg2rad(360.f), deg180 = deg2rad(180.f);
  double CPrimeProduct = c.C1p * c.C2p;
  double h1, h2;
  if ((double)B1 == 0.0 && c.a1p == 0.0) h1 = 0.0;
  else {
    h1 = nqm::nq_atan2((double)B1, c.a1p);
    if (h1 < 0) h1 += (double)deg360;
  }
  if ((double)B2 == 0.0 && c.a2p == 0.0) h2 = 0.0;
  else {
    h2 = nqm::nq_atan2((double)B2, c.a2p);
    if (h2 < 0) h2 += (double)deg360;
  }
  double dh;
  if (CPrimeProduct == 0.0) dh = 0;
  else {
    dh = h2 - h1;
    if (dh < (double)(-deg180)) dh += (double)deg360;
    else if (dh > (double)deg180) dh -= (double)deg360;
  }
  double deltaHPrime = 2.0 * nqm::sqrt_(CPrimeProduct) * nqm::nq_sin(dh / 2.0);
  double hsum = h1 + h2, bh;
  if (CPrimeProduct == 0.0) bh = hsum;
  else {
    if (nqm::fabs_(h1 - h2) <= (double)deg180) bh = hsum / 2.0;
    else if (hsum < (double)deg360) bh = (hsum + (double)deg360) / 2.0;
    else bh = (hsum - (double)deg360) / 2.0;
  }
  *barhPrime = bh;
  *barCPrime = (c.C1p + c.C2p) / 2.0;
  double T = 1.0 - (0.17 * nqm::nq_cos(bh - (double)deg2rad(30.f))) + (0.24 * nqm::nq_cos(2.0 * bh)) +
             (0.32 * nqm::nq_cos((3.0 * bh) + (double)deg2rad(6.f))) - (0.20 * nqm::nq_cos((4.0 * bh) - (double)deg2rad(63.f)));
  double S_H = 1 + ((double)0.015f * *barCPrime * T);
  return (float)(deltaHPrime / (1.0 * S_H));
}

// R_T term (CL:187-194)
NQ_HD float ciede_RT(double barCPrime, double barhPrime, float Cterm, float Hterm) {
  double q = (barhPrime - (double)deg2rad(275.f)) / (double)deg2rad(25.f);
  double deltaTheta = (double)deg2rad(30.f) * nqm::nq_exp(-(q * q));
  double p7 = nqm::nq_pow(barCPrime, 7.0);
  double R_C = 2.0 * nqm::sqrt_(p7 / (p7 + 6103515625.0));
  double RT = (-nqm::nq_sin(2.0 * deltaTheta)) * R_C;
  return (float)(RT * (double)Cterm * (double)Hterm);
}

// ---- luminance / chroma deltas (CL:215-238) -----------------------------------------------------
NQ_HD double color_y(uint32_t c, const double* lut) {
  return lut[c_red(c)] * 0.2126 + lut[c_green(c)] * 0.7152 + lut[c_blue(c)] * 0.0722;
}
NQ_HD double y_diff(uint32_t c1, uint32_t c2, const double* lut) {
  double y = color_y(c1, lut), y2 = color_y(c2, lut);
  return nqm::fabs_(y2 - y) * 100;
}
NQ_HD double color_u(uint32_t c) { return -0.09991 * c_red(c) - 0.33609 * c_green(c) + 0.436 * c_blue(c); }
NQ_HD double u_diff(uint32_t c1, uint32_t c2) { return nqm::fabs_(color_u(c2) - color_u(c1)); }

// ---- BlueNoise.diffuse (BN:180-197) ---------------------------------------------------------------
NQ_HD uint32_t bn_diffuse(uint32_t pixel, uint32_t qPixel, float weight, float strength, int x, int y, const signed char* bn) {
  int r = c_red(pixel), g = c_green(pixel), b = c_blue(pixel), a = c_alpha(pixel);
  float adj = ((float)bn[(x & 63) | (y & 63) << 6] + 0.5f) / 127.5f;
  adj += ((float)((x + y) & 1) - 0.5f) * strength / 8.f;
  adj *= weight;
  r = j2i(dmin(255.0, dmax((double)((float)r + (adj * (float)(r - c_red(qPixel)))), 0.0)));
  g = j2i(dmin(255.0, dmax((double)((float)g + (adj * (float)(g - c_green(qPixel)))), 0.0)));
  b = j2i(dmin(255.0, dmax((double)((float)b + (adj * (float)(b - c_blue(qPixel)))), 0.0)));
  a = j2i(dmin(255.0, dmax((double)((float)a + (adj * (float)(a - c_alpha(qPixel)))), 0.0)));
  return c_argb(a, r, g, b);
}

// ---- java.util.Random (48-bit LCG) behind PnnLABQuantizer.closestColorIndex (PL:22,467) ------------
struct JRandom {
  unsigned long long seed;
  NQ_HD void set_seed(unsigned long long s) { seed = (s ^ 0x5DEECE66DULL) & ((1ULL << 48) - 1); }
  NQ_HD int next31() {
    seed = (seed * 0x5DEECE66DULL + 0xBULL) & ((1ULL << 48) - 1);
    return (int)(seed >> 17);
  }
  // nextInt(bound) for a bound that is not a power of two
  NQ_HD int next_int(int bound) {
    int u = next31(), r = u % bound;
    while ((int)((unsigned)(u - r) + (unsigned)(bound - 1)) < 0) { u = next31(); r = u % bound; }
    return r;
  }
};

}  // namespace nq
