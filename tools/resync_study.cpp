// resync_study.cpp -- CPU experiment for SURVEY.md section 8(f) rank 1 (speculative segment-parallel
// Gilbert dither). TEST/STUDY INFRASTRUCTURE: builds on the oracle, never part of the product.
//
// Question: if the error diffusion of GilbertCurve.diffusePixel (GC:187-280) is restarted at curve
// position p0 with an EMPTY error queue (all boxes zero) instead of the true queue, after how many
// pixels does the queue become bit-identical to the one of the true run again? From that point on a
// speculative segment reproduces the sequential result exactly, so its length bounds the warm-up a
// segment-parallel kernel needs.
//
// Method: the true run records, per pixel, a hash of the queue after the pixel and every palette lookup
// (colour asked, index answered). A speculative run from p0 replays the recorded answer whenever it asks
// for the same colour at the same pixel (the memo / java.util.Random state a real implementation would
// have to carry is thereby idealised away) and falls back to the live quantizer otherwise.
#include <cstdint>
#include <cstring>
#include <cmath>
#include <vector>
#include <deque>
#include <unordered_map>
#include <array>
#include <string>
#include <algorithm>
#include <stdexcept>
#define private public
#define protected public
#define class struct
#include "../oracle/nq_oracle.cpp"
#undef private
#undef protected
#undef class

namespace {

struct StudyCfg {
  int nstarts = 16, maxlen = 200000;
  std::vector<int> lengths;      // per start: pixels until the queue matched again, -1 = not within maxlen
  std::vector<int> starts;
  long long lookupsDiffused = 0; // lookups of the true run whose colour differs from a zero-error lookup (informational)
  int DM = 0, sorted = 0, ditherMax = 0;
};
StudyCfg* g_cfg = nullptr;

struct Call { int32_t c; short r; };

struct Recorder : Ditherable {
  Ditherable& inner;
  std::vector<uint32_t> first;   // index into calls of the first call of pixel n (by curve position)
  std::vector<Call> calls;
  int cur = 0;
  bool replay = false;
  long long misses = 0;
  explicit Recorder(Ditherable& in) : inner(in) {}
  int getColorIndex(int32_t c) override { return inner.getColorIndex(c); }
  short nearestColorIndex(const std::vector<int32_t>& palette, int32_t c, int pos) override {
    if (!replay) {
      short r = inner.nearestColorIndex(palette, c, pos);
      calls.push_back(Call{c, r});
      return r;
    }
    for (uint32_t k = first[cur]; k < first[cur + 1]; ++k)
      if (calls[k].c == c) return calls[k].r;
    ++misses;
    return inner.nearestColorIndex(palette, c, pos);
  }
};

uint64_t mix(uint64_t h, uint64_t v) {
  h ^= v + 0x9E3779B97F4A7C15ULL + (h << 6) + (h >> 2);
  h *= 0xBF58476D1CE4E5B9ULL;
  return h ^ (h >> 29);
}
uint64_t state_hash(const GilbertCurve& g) {
  uint64_t h = 0x1234567;
  const size_t n = g.qSize();
  for (size_t i = 0; i < n; ++i) {
    const auto& b = g.qAt(i);
    uint32_t u[4];
    memcpy(u, b.p, 16);
    h = mix(h, ((uint64_t)u[0] << 32) | u[1]);
    h = mix(h, ((uint64_t)u[2] << 32) | u[3]);
    if (g.sortedByYDiff) { uint64_t y; memcpy(&y, &b.yDiff, 8); h = mix(h, y); }
  }
  return mix(h, n);
}

std::vector<int32_t> run_study(int width, int height, const std::vector<int32_t>& cPixels, std::vector<int32_t>& palette,
                               Ditherable& inner, const std::vector<float>* sal, double weight, bool dither) {
  StudyCfg& cfg = *g_cfg;
  const int npix = width * height;
  OrderOnly oo;
  oo.width = width;
  oo.out.reserve(npix);
  if (width >= height) oo.gen(0, 0, width, 0, 0, height); else oo.gen(0, 0, 0, height, width, 0);
  const std::vector<uint32_t>& order = oo.out;

  Recorder rec(inner);
  std::vector<int32_t> qPixels(npix, 0);
  std::vector<uint64_t> H(npix);
  GilbertCurve gc(width, height, cPixels, palette, qPixels, rec, sal, weight, dither);
  cfg.DM = gc.DITHER_MAX; cfg.sorted = gc.sortedByYDiff; cfg.ditherMax = gc.ditherMax;
  if (!gc.sortedByYDiff) gc.initWeights(gc.DITHER_MAX);
  rec.first.resize(npix + 1);
  for (int n = 0; n < npix; ++n) {
    rec.first[n] = (uint32_t)rec.calls.size();
    const int bidx = (int)order[n];
    gc.diffusePixel(bidx % width, bidx / width);
    H[n] = state_hash(gc);
  }
  rec.first[npix] = (uint32_t)rec.calls.size();

  // speculative restarts
  rec.replay = true;
  std::vector<int32_t> q2(npix, 0);
  cfg.lengths.clear(); cfg.starts.clear();
  for (int s = 0; s < cfg.nstarts; ++s) {
    const int p0 = (int)((long long)(s + 1) * npix / (cfg.nstarts + 1));
    GilbertCurve g2(width, height, cPixels, palette, q2, rec, sal, weight, dither);
    if (!g2.sortedByYDiff) g2.initWeights(g2.DITHER_MAX);
    else {
      g2.initWeights(7);                                  // steady state of the PriorityQueue mode: weights of length 7,
      while (g2.qSize() < 15) g2.qAdd(decltype(g2.heap)::value_type());   // 15 boxes (GC:231-234, 345)
    }
    int len = -1;
    const int end = std::min(npix, p0 + cfg.maxlen);
    for (int n = p0; n < end; ++n) {
      rec.cur = n;
      const int bidx = (int)order[n];
      q2[bidx] = 0;
      g2.diffusePixel(bidx % width, bidx / width);
      if (state_hash(g2) == H[n] && q2[bidx] == qPixels[bidx]) { len = n - p0 + 1; break; }
    }
    cfg.starts.push_back(p0);
    cfg.lengths.push_back(len);
  }
  return qPixels;
}

struct StudyRgb : PnnQuantizer {
  using PnnQuantizer::PnnQuantizer;
  std::vector<int32_t> ditherImage(const std::vector<int32_t>& cPixels, std::vector<int32_t>& palette, int width, int height, bool dither) override {
    RgbDitherable ditherable(*this, dither);
    if (hasSemiTransparency) weight *= -1;
    return run_study(width, height, cPixels, palette, ditherable, nullptr, weight, dither);
  }
};
struct StudyLab : PnnLABQuantizer {
  using PnnLABQuantizer::PnnLABQuantizer;
  std::vector<int32_t> ditherImage(const std::vector<int32_t>& cPixels, std::vector<int32_t>& palette, int width, int height, bool dither) override {
    LabDitherable ditherable(*this);
    random.setSeed(rngSeed);
    if (hasSemiTransparency) weight *= -1;
    if (dither && !hasSaliencies && (palette.size() <= 256 || weight > .99)) {   // PL:499-508
      saliencies.assign(pixels.size(), 0.f);
      hasSaliencies = true;
      float saliencyBase = .1f;
      for (size_t i = 0; i < pixels.size(); ++i) {
        const Lab& lab1 = getLab(pixels[i]);
        saliencies[i] = saliencyBase + (1 - saliencyBase) * lab1.L / 100.f * lab1.alpha / 255.f;
      }
    }
    return run_study(width, height, cPixels, palette, ditherable, hasSaliencies ? &saliencies : nullptr, weight, dither);
  }
};

}  // namespace

extern "C" int nqs_resync_study(int kind, const uint32_t* argb, int w, int h, int nmax, int dither, uint64_t seed, int nstarts, int maxlen,
                                int* starts, int* lengths, int* info /* DM, sorted, ditherMax, paletteLen */) {
  StudyCfg cfg;
  cfg.nstarts = nstarts; cfg.maxlen = maxlen;
  g_cfg = &cfg;
  M.mode = 0;
  try {
    int plen = 0;
    if (kind == 0) { StudyRgb q(argb, w, h); q.rngSeed = seed; q.convert(nmax, dither != 0); plen = (int)q.palette_out.size(); }
    else { StudyLab q(argb, w, h); q.rngSeed = seed; q.convert(nmax, dither != 0); plen = (int)q.palette_out.size(); }
    for (int i = 0; i < (int)cfg.lengths.size(); ++i) { starts[i] = cfg.starts[i]; lengths[i] = cfg.lengths[i]; }
    info[0] = cfg.DM; info[1] = cfg.sorted; info[2] = cfg.ditherMax; info[3] = plen;
  } catch (const std::exception& e) {
    g_cfg = nullptr;
    return -1;
  }
  g_cfg = nullptr;
  return (int)cfg.lengths.size();
}
