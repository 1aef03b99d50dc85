// spec_host_harness.cpp -- TEST INFRASTRUCTURE. Compiles the stage bodies of
// nquant_android_b200/csrc/nq_dither_spec.cuh (scalar host/device functions) with g++ and runs the whole
// speculative segment-parallel dither on the CPU, image constants taken from the oracle, result compared
// with the oracle's sequential GilbertCurve (GC:187-280). The CUDA kernels wrap the same functions, so this
// checks their arithmetic and the validation logic without a GPU (tests/test_spec_dither_host.py).
#include <cstdint>
#include <cstring>
#include <cmath>
#include <cstdio>
#include <cstdlib>
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
struct uint2 { unsigned x, y; };   // CUDA vector type named by nq_types.h
// ---- the kernels of nq_dither_spec.cuh, run thread by thread on the CPU (NQS_EMULATE) -------------------------------------
#define NQS_EMULATE 1
struct dim3 { unsigned x, y, z; dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
static thread_local dim3 blockIdx, threadIdx, blockDim, gridDim;
#define __global__ static
#define __launch_bounds__(x)
static inline int atomicMin(int* p, int v) { int o = *p; if (v < o) *p = v; return o; }
static inline int atomicAdd(int* p, int v) { int o = *p; *p += v; return o; }
namespace nq { static double g_gammaLut[256]; static signed char g_blueNoise[4096]; }   // device globals of nq_hist.cuh
#include "../nquant_android_b200/csrc/nq_dither_spec.cuh"

namespace {

struct HarnessOut {
  long long eligible = 0, exact = 0, rounds = 0, anomaly = 0, nseg = 0, segRuns = 0, slowPixels = 0, notes = 0, mismatches = 0, rejected = 0, patches = 0, redos = 0;
};
struct HarnessCfg { int seg = 4096, warm = 1024, useCells = 1; HarnessOut out; };
HarnessCfg* g_h = nullptr;

// one image of a batch for the kernel-level emulation (nqs_spec_batch_*)
struct Collected {
  NqImage I;
  std::vector<uint32_t> in, ref;
  std::vector<unsigned char> cells;
};
std::vector<Collected> g_batch;
bool g_collect = false;

// host copy of k_build_cells (nq_dither.cuh): candidate lists per 5-5-5 RGB cell
void build_cells(const nq::spec::SpecConst& C, std::vector<unsigned char>& cells) {
  cells.assign((size_t)32768 * 32, 0);
  for (int cell = 0; cell < 32768; ++cell) {
    const int r0 = (cell >> 10) << 3, g0 = ((cell >> 5) & 31) << 3, b0 = (cell & 31) << 3;
    int h0 = 0x7fffffff, h1 = 0x7fffffff;
    for (int k = 0; k < C.plen; ++k) {
      const uint32_t pc = C.pal[k];
      const int pr = nq::c_red(pc), pg = nq::c_green(pc), pb = nq::c_blue(pc);
      const double hi = C.Tr[std::max(std::abs(pr - r0), std::abs(pr - r0 - 7))] + C.Tg[std::max(std::abs(pg - g0), std::abs(pg - g0 - 7))] +
                        C.Tb[std::max(std::abs(pb - b0), std::abs(pb - b0 - 7))];
      const int d = nq::j2i(hi * (1.0 + 1e-12) + 1e-9);
      if (d < h0) { h1 = h0; h0 = d; } else if (d < h1) h1 = d;
    }
    unsigned char* o = &cells[(size_t)cell * 32];
    int cnt = 0;
    for (int k = 0; k < C.plen; ++k) {
      const uint32_t pc = C.pal[k];
      const int pr = nq::c_red(pc), pg = nq::c_green(pc), pb = nq::c_blue(pc);
      const double lo = C.Tr[std::max(0, std::max(r0 - pr, pr - r0 - 7))] + C.Tg[std::max(0, std::max(g0 - pg, pg - g0 - 7))] +
                        C.Tb[std::max(0, std::max(b0 - pb, pb - b0 - 7))];
      const int d = nq::j2i(lo * (1.0 - 1e-12) - 1e-9);
      if (d <= h1) { ++cnt; if (cnt <= 31) o[cnt] = (unsigned char)k; }
    }
    o[0] = cnt > 31 ? 255 : (unsigned char)cnt;
  }
}

void run_spec(PnnLABQuantizer& q, const std::vector<int32_t>& cPixels, const std::vector<int32_t>& palette, int width, int height,
              bool dither, double weightSigned, int nMaxColors, const std::vector<int32_t>& reference) {
  using namespace nq::spec;
  HarnessCfg& H = *g_h;
  HarnessOut& R = H.out;
  const int npix = width * height;
  std::vector<int32_t> scratch(npix, 0);
  PnnLABQuantizer::LabDitherable dummy(q);
  GilbertCurve gc(width, height, cPixels, palette, scratch, dummy, q.hasSaliencies ? &q.saliencies : nullptr, weightSigned, dither);
  const int plen = (int)palette.size();
  // same eligibility rule as the device (k_spec_setup)
  const int acceptedDiff = std::max(2, plen - gc.margin);
  const bool eligible = dither && q.hasSaliencies && !gc.sortedByYDiff && !gc.hasAlpha && plen > 64 && 2 * acceptedDiff > 101 && q.m_transparentPixelIndex < 0;
  R.eligible = eligible;
  if (!eligible) return;

  static SpecConst C;
  memset(&C, 0, sizeof(C));
  C.plen = plen; C.margin = gc.margin; C.thresold = gc.thresold; C.DM = gc.DITHER_MAX; C.ditherMax = gc.ditherMax;
  C.width = width; C.npix = npix;
  C.isNano = q.isNano; C.hasTrans = q.m_transparentPixelIndex >= 0; C.salReplaced = nMaxColors < 128 && nMaxColors > 2;
  C.seg = H.seg; C.warm = H.warm; C.nseg = (npix + H.seg - 1) / H.seg;
  C.transColor = (uint32_t)q.m_transparentColor;
  C.gWeight = gc.weight; C.PR = q.PR; C.PG = q.PG; C.PB = q.PB; C.ratio = q.ratio;
  C.beta = gc.beta;
  C.seed0 = (q.rngSeed ^ 0x5DEECE66DULL) & ((1ULL << 48) - 1);
  gc.initWeights(gc.DITHER_MAX);
  for (int k = 0; k < C.DM; ++k) C.w[k] = gc.weights[k];
  for (int i = 0; i < plen; ++i) C.pal[i] = (uint32_t)palette[i];
  std::vector<double> lut(256);
  for (int v = 0; v < 256; ++v) lut[v] = nq::gamma_to_linear(v);
  fill_tables(C, lut.data());
  {   // the generator steps nextInt(32767) rejects (k_spec_rejects / k_spec_rejsort on the device)
    const unsigned long long MASK = (1ULL << 48) - 1;
    unsigned long long st = C.seed0;
    C.nrej = 0;
    for (unsigned long long t = 1; t <= (unsigned long long)npix + NQS_MAXREJ + 1; ++t) {
      st = (st * 0x5DEECE66DULL + 0xBULL) & MASK;
      if (lcg_rejects(st) && C.nrej < NQS_MAXREJ) C.rej[C.nrej++] = (unsigned)t;
    }
  }
  C.opaque = 1;
  for (int i = 0; i < npix; ++i) C.opaque &= ((uint32_t)cPixels[i] >> 24) == 0xFFu;
  for (int i = 0; i < plen; ++i) C.opaque &= (C.pal[i] >> 24) >= 0xFEu;
  std::vector<float> tanhTab(512, 0.f);   // the table k_spec_tables builds on the device (shape_tanh)
  for (int t = 0; t < 511; ++t) tanhTab[t] = tanh_f((double)((float)(t - 255) / 255.f * 20.f));

  OrderOnly oo;
  oo.width = width;
  if (width >= height) oo.gen(0, 0, width, 0, 0, height); else oo.gen(0, 0, 0, height, width, 0);
  std::vector<uint32_t> order(npix);
  for (int n = 0; n < npix; ++n) order[n] = (oo.out[n] % width) | ((oo.out[n] / width) << 16);

  std::vector<uint32_t> in(npix), out(npix, 0), cpx(npix), ccol(npix), ck0(npix), ck1(npix), cdraw(npix + 1);
  for (int i = 0; i < npix; ++i) in[i] = (uint32_t)cPixels[i];
  std::vector<unsigned short> cq(npix), memo(65536, 0xFFFF), slowVal(65536, 0);
  std::vector<unsigned char> cflag(npix), cells;
  std::vector<int> firstPos(65536, NQS_NOPOS), slowPos(65536, NQS_NOPOS), state(16, 0);
  std::vector<SpecRec> rec((size_t)C.nseg * C.seg);
  std::vector<SpecSeg> segs(C.nseg);
  memset(segs.data(), 0, sizeof(SpecSeg) * segs.size());
  for (auto& s : segs) { s.dirty = 1; s.warmMul = 1; s.mispos = -1; }
  segs[0].exact = 1;
  if (getenv("NQ_SPEC_FORCECHAIN")) {   // debugging aid: the whole image as ONE sequential chain from the first pixel
    const int k = std::min(C.nseg, atoi(getenv("NQ_SPEC_FORCECHAIN")));
    segs[0].chain = k;
    for (int t = 1; t < k; ++t) { segs[t].chained = 1; segs[t].exact = 1; }
  }
  if (H.useCells) build_cells(C, cells);
  if (g_collect) {   // what k_dither_setup leaves in NqImage for this image (the fields k_spec_setup reads)
    Collected B;
    memset(&B.I, 0, sizeof(B.I));
    NqImage& I = B.I;
    I.kind = NQ_KIND_LAB; I.width = width; I.height = height; I.npix = npix; I.nmax = nMaxColors; I.dither = 1;
    I.seed = q.rngSeed; I.transIdx = q.m_transparentPixelIndex; I.hasSemi = q.hasSemiTransparency; I.transColor = (uint32_t)q.m_transparentColor;
    I.isNano = q.isNano; I.PR = q.PR; I.PG = q.PG; I.PB = q.PB; I.PA = q.PA; I.ratioMerge = q.ratio;
    I.paletteLen = plen;
    for (int i = 0; i < plen; ++i) I.palette[i] = (uint32_t)palette[i];
    I.gMargin = gc.margin; I.gThresold = gc.thresold; I.gDitherMaxQ = gc.DITHER_MAX; I.gDitherMax = gc.ditherMax;
    I.gSorted = gc.sortedByYDiff; I.gHasAlpha = gc.hasAlpha; I.gUseSal = q.hasSaliencies; I.gBeta = gc.beta; I.gWeight = gc.weight;
    for (int k = 0; k < C.DM; ++k) I.gWeights[k] = gc.weights[k];
    for (int i = 0; i < npix; ++i) I.nonOpaque += ((uint32_t)cPixels[i] >> 24) != 0xFFu;
    B.in.resize(npix); B.ref.resize(npix);
    for (int i = 0; i < npix; ++i) { B.in[i] = (uint32_t)cPixels[i]; B.ref[i] = (uint32_t)reference[i]; }
    build_cells(C, B.cells);
    g_batch.push_back(std::move(B));
    return;
  }
  static const signed char bn[4096] = NQ_BLUE_NOISE_INIT;

  SpecWork W;
  W.order = order.data(); W.in = in.data(); W.out = out.data(); W.cpx = cpx.data(); W.ccol = ccol.data(); W.ck0 = ck0.data(); W.ck1 = ck1.data();
  W.cq = cq.data(); W.cflag = cflag.data(); W.cdraw = cdraw.data(); W.firstPos = firstPos.data(); W.memo = memo.data();
  W.slowPos = slowPos.data(); W.slowVal = slowVal.data(); W.cells = H.useCells ? cells.data() : nullptr; W.lut = lut.data(); W.bn = bn;
  W.segs = segs.data(); W.state = state.data(); W.rec = rec.data();
  std::vector<int> patchList(2 * NQS_MAXPATCH, 0);
  W.patch = patchList.data();

  for (int n = 0; n < npix; ++n) stage_pre(C, W, n);                                     // stage 1
  { uint32_t d = 0; for (int n = 0; n < npix; ++n) { cdraw[n] = d; d += (cflag[n] & NQS_F_DRAW) ? 1u : 0u; } cdraw[npix] = d; }   // stage 2
  for (int n = 0; n < npix; ++n) {                                                       // stage 3
    int key;
    if (!stage_resolve(C, W, n, &key)) { ++R.rejected; state[1] = 1; }
    if (key >= 0 && n < firstPos[key]) firstPos[key] = n;
    if (!(cflag[n] & NQS_F_PRE)) ++R.slowPixels;
    if (cflag[n] & NQS_F_RISK) ++state[7];
    if (!(cflag[n] & NQS_F_PRE)) ++state[8];
  }
  stage_gate(C, W);
  for (int key = 0; key < 65536; ++key) stage_memo(C, W, key, -1);                       // stage 4
  for (int n = 0; n < npix; ++n) stage_fill(C, W, n);                                    // stage 5
  for (int n = 0; n < npix; ++n) stage_pack(C, W, n);                                    // stage 5b
  R.nseg = C.nseg;
  int open = C.nseg;
  while (open > 0 && !state[1] && R.rounds < 100000) {
    ++R.rounds;
    for (int s = 0; s < C.nseg; ++s) { if (!segs[s].done && segs[s].dirty) ++R.segRuns; stage_run(C, W, s, tanhTab.data()); }   // stage 6
    for (int s = 0; s < C.nseg; ++s) stage_compare(C, W, s);                             // stage 6b
    open = stage_validate(C, W);                                                         // stage 7
    if (getenv("NQ_SPEC_DEBUG")) fprintf(stderr, "round %lld: validated up to segment %d, open %d, patch %d, redo from %d (rekey %d)\n", R.rounds, state[0], open, state[2], state[5] - 1, state[10] - 1);
    if (getenv("NQ_SPEC_DEBUG2") && open > 0 && !state[1]) {
      const int s = state[0];
      int diff = 0;
      if (s > 0) for (int k = 0; k < C.DM; ++k) for (int j = 0; j < 4; ++j) diff += segs[s].qwarm[k][j] != segs[s - 1].qout[k][j];
      fprintf(stderr, "round %lld: stopped at segment %d (exact %d, notes %d, %d queue floats differ from the predecessor's)\n", R.rounds, s, segs[s].exact, segs[s].nnotes, diff);
      if (s > 0 && R.rounds < 3) for (int k = 0; k < C.DM; ++k) fprintf(stderr, "  box %2d warm (%g %g %g %g) prev (%g %g %g %g)\n", k, segs[s].qwarm[k][0], segs[s].qwarm[k][1], segs[s].qwarm[k][2], segs[s].qwarm[k][3],
          segs[s - 1].qout[k][0], segs[s - 1].qout[k][1], segs[s - 1].qout[k][2], segs[s - 1].qout[k][3]);
    }
    const bool repack = state[2] || state[5];
    if (state[5] && !state[1]) {                                                         // draw misprediction: re-resolve behind it
      ++R.redos;
      const int from = state[5] - 1;
      for (int n = 0; n < npix; ++n) stage_adopt(W, n);
      { uint32_t d = 0; for (int n = 0; n < npix; ++n) { cdraw[n] = d; d += (cflag[n] & NQS_F_DRAW) ? 1u : 0u; } cdraw[npix] = d; }
      for (int key = 0; key < 65536; ++key) stage_rekey(W, key, from);
      for (int n = from + 1; n < npix; ++n) {
        int key;
        if (!stage_resolve(C, W, n, &key)) { ++R.rejected; state[1] = 1; }
        if (key >= 0 && n < firstPos[key]) firstPos[key] = n;
      }
      for (int key = 0; key < 65536; ++key) stage_memo(C, W, key, from);
      for (int n = from + 1; n < npix; ++n) stage_fill(C, W, n);
      state[5] = 0;
    }
    if (state[2]) { ++R.patches; for (int n = 0; n < npix; ++n) stage_patch(C, W, n); state[2] = 0; }
    if (repack) for (int n = 0; n < npix; ++n) stage_pack(C, W, n);
  }
  R.anomaly = state[1];
  if (getenv("NQ_SPEC_DEBUG")) fprintf(stderr, "risk pixels %d, error-dependent lookups %d of %d\n", state[7], state[8], npix);
  for (auto& s : segs) R.notes += std::min(s.nnotes, NQS_NOTES);
  if (!state[1] && getenv("NQ_SPEC_DEBUG") && !q.trace.draw_bidx.empty()) {
    // the oracle's draws per pixel against the flags the path ended with
    std::vector<unsigned char> truth(npix, 0);
    for (int32_t b : q.trace.draw_bidx) truth[b] = 1;
    int shown = 0;
    unsigned d = 0;
    for (int n = 0; n < npix && shown < 8; ++n) {
      const int bidx = (int)(order[n] & 0xFFFF) + (int)(order[n] >> 16) * width;
      const int mine = (cflag[n] & NQS_F_DRAW) ? 1 : 0;
      if (mine != truth[bidx] || cdraw[n] != d) {
        fprintf(stderr, "draw flag / prefix differs at curve position %d (segment %d offset %d): flag %u (truth draws %d), cdraw %u (truth %u)\n", n, n / C.seg, n % C.seg, (unsigned)cflag[n], truth[bidx], cdraw[n], d);
        ++shown;
      }
      d += truth[bidx];
    }
  }
  if (!state[1] && getenv("NQ_SPEC_DEBUG")) {
    int shown = 0;
    for (int n = 0; n < npix && shown < 6; ++n) {
      const int bidx = (int)(order[n] & 0xFFFF) + (int)(order[n] >> 16) * width;
      if (out[bidx] != (uint32_t)reference[bidx]) {
        const int sg = n / C.seg;
        fprintf(stderr, "first mismatch at curve position %d (segment %d, offset %d): flag %u cq %d cdraw %u; seg exact %d chain %d chained %d dev %d draws %d nslow %d notes %d\n", n, sg, n % C.seg,
                (unsigned)cflag[n], (int)cq[n], cdraw[n], segs[sg].exact, segs[sg].chain, segs[sg].chained, segs[sg].dev, segs[sg].draws, segs[sg].nslow, segs[sg].nnotes);
        int mi = -1, ri = -1;
        for (int k = 0; k < plen; ++k) { if (C.pal[k] == out[bidx]) mi = k; if (C.pal[k] == (uint32_t)reference[bidx]) ri = k; }
        const SpecRec& rr = rec[rec_index(C, n)];
        fprintf(stderr, "   ours %08x (palette %d), reference %08x (palette %d); record qf %08x; ck0 %u/%u ck1 %u/%u ccol %08x\n", out[bidx], mi, (uint32_t)reference[bidx], ri, rr.qf,
                ck0[n] >> 8, ck0[n] & 255, ck1[n] >> 8, ck1[n] & 255, ccol[n]);
        ++shown;
        n = (sg + 1) * C.seg - 1;
      }
    }
  }
  if (!state[1]) {
    for (int i = 0; i < npix; ++i) R.mismatches += out[i] != (uint32_t)reference[i];
    R.exact = R.mismatches == 0;
  }
}

struct HostLab : PnnLABQuantizer {
  using PnnLABQuantizer::PnnLABQuantizer;
  int nMax = 0;
  std::vector<int32_t> ditherImage(const std::vector<int32_t>& cPixels, std::vector<int32_t>& palette, int width, int height, bool dither) override {
    const double w0 = weight;
    std::vector<int32_t> ref = PnnLABQuantizer::ditherImage(cPixels, palette, width, height, dither);
    weight = w0;
    const double ws = hasSemiTransparency ? -weight : weight;
    run_spec(*this, cPixels, palette, width, height, dither, ws, nMax, ref);
    return ref;
  }
};

}  // namespace

extern "C" int nqs_spec_host(const uint32_t* argb, int w, int h, int nmax, int dither, uint64_t seed, int seg, int warm, int useCells,
                             long long* out /* 12 values */) {
  HarnessCfg cfg;
  cfg.seg = seg; cfg.warm = warm; cfg.useCells = useCells;
  g_h = &cfg;
  M.mode = 0;
  try {
    HostLab q(argb, w, h);
    q.rngSeed = seed; q.nMax = nmax;
    q.trace.keep_draws = getenv("NQ_SPEC_DEBUG") != nullptr;
    q.convert(nmax, dither != 0);
  } catch (const std::exception& e) {
    fprintf(stderr, "spec host harness: %s\n", e.what());
    g_h = nullptr;
    return -1;
  }
  const HarnessOut& R = cfg.out;
  long long v[12] = {R.eligible, R.exact, R.rounds, R.anomaly, R.nseg, R.segRuns, R.slowPixels, R.notes, R.mismatches, R.rejected, R.patches, R.redos};
  memcpy(out, v, sizeof(v));
  g_h = nullptr;
  return 0;
}

// ---- kernel-level emulation of a batch: k_spec_setup, spec_bind, spec_drive with the kernels run thread by thread -------
namespace {
struct EmuBackend {
  long long launches = 0;
  template <class... P, class... A>
  void launch(void (*kernel)(P...), dim3 grid, int block, A... args) {
    gridDim = grid; blockDim = dim3((unsigned)block);
    for (unsigned y = 0; y < grid.y; ++y)
      for (unsigned x = 0; x < grid.x; ++x) {
        blockIdx = dim3(x, y);
        for (unsigned t = 0; t < (unsigned)block; ++t) { threadIdx = dim3(t); kernel(args...); }
      }
    ++launches;
  }
  template <class... P, class... A>
  void launch_run(void (*kernel)(P...), dim3 grid, int block, A... args) { launch(kernel, grid, block, args...); }
  void write_ints(int* dev, const int* host, int n) { memcpy(dev, host, sizeof(int) * (size_t)n); }
  void read_ints(int* host, const int* dev, int n) { memcpy(host, dev, sizeof(int) * (size_t)n); }
  void lap(const char*) {}
  void note(int, int, int, int, int) {}
  void done_prefix(int) {}
};
}  // namespace

extern "C" void nqs_spec_batch_begin() { g_batch.clear(); }
extern "C" int nqs_spec_batch_add(const uint32_t* argb, int w, int h, int nmax, uint64_t seed, int seg, int warm) {
  HarnessCfg cfg;
  cfg.seg = seg; cfg.warm = warm; cfg.useCells = 1;
  g_h = &cfg; g_collect = true;
  M.mode = 0;
  int rc = 0;
  try {
    HostLab q(argb, w, h);
    q.rngSeed = seed; q.nMax = nmax;
    q.convert(nmax, true);
  } catch (const std::exception& e) {
    fprintf(stderr, "spec host harness: %s\n", e.what());
    rc = -1;
  }
  g_h = nullptr; g_collect = false;
  return rc;
}
// out: images, eligible, completed, handed back, rounds, patches, re-resolves, images with wrong pixels, launches
extern "C" int nqs_spec_batch_run(int seg, int warm, int wave, long long* out /* 9 values */) {
  using namespace nq::spec;
  const int n = (int)g_batch.size();
  if (!n) return -1;
  const int npix = g_batch[0].I.npix, width = g_batch[0].I.width, height = g_batch[0].I.height;
  for (int v = 0; v < 256; ++v) nq::g_gammaLut[v] = nq::gamma_to_linear(v);
  static const signed char bn[4096] = NQ_BLUE_NOISE_INIT;
  memcpy(nq::g_blueNoise, bn, 4096);
  OrderOnly oo;
  oo.width = width;
  if (width >= height) oo.gen(0, 0, width, 0, 0, height); else oo.gen(0, 0, 0, height, width, 0);
  std::vector<uint32_t> order(npix);
  for (int i = 0; i < npix; ++i) order[i] = (oo.out[i] % width) | ((oo.out[i] / width) << 16);
  std::vector<NqImage> imgs(n);
  std::vector<NqSlot> slots(n);
  std::vector<std::vector<uint32_t>> outs(n, std::vector<uint32_t>(npix, 0));
  for (int i = 0; i < n; ++i) {
    imgs[i] = g_batch[i].I;
    memset(&slots[i], 0, sizeof(NqSlot));
    slots[i].in = g_batch[i].in.data(); slots[i].out = outs[i].data(); slots[i].cells = g_batch[i].cells.data();
  }
  const SpecLayout L = spec_layout(npix, seg);
  if (wave < 1 || wave > n) wave = n;
  std::vector<unsigned char> buf(L.perSlot * (size_t)wave);
  std::vector<SpecImage> sp(n);
  memset(sp.data(), 0, sizeof(SpecImage) * (size_t)n);
  std::vector<SpecWork> pool(wave);
  spec_bind_pool(pool.data(), wave, buf.data(), L);
  std::vector<int> elig(n, 0), ints(4 * wave + 4, 0), handed(n, 0);
  std::vector<float> tanhTab(512, 0.f);
  for (int t = 0; t < 511; ++t) tanhTab[t] = tanh_f((double)((float)(t - 255) / 255.f * 20.f));
  EmuBackend be;
  be.launch(k_spec_setup, dim3((n + 63) / 64), 64, imgs.data(), (const NqSlot*)slots.data(), sp.data(), (const uint32_t*)order.data(), n, seg, warm, elig.data());
  SpecStats st;
  spec_drive(be, imgs.data(), sp.data(), (const SpecWork*)pool.data(), (const int*)elig.data(), n, npix, seg, wave, ints.data(), (const float*)tanhTab.data(), 148, &st, handed.data());
  long long eligible = 0, wrong = 0;
  for (int i = 0; i < n; ++i) {
    eligible += (elig[i] & 255) == 1;
    if (imgs[i].specDone == 1) {
      long long bad = 0;
      for (int k = 0; k < npix; ++k) bad += outs[i][k] != g_batch[i].ref[k];
      wrong += bad != 0;
    }
  }
  long long v[9] = {n, eligible, (long long)st.done, (long long)st.handedBack, (long long)st.rounds, (long long)st.patches, (long long)st.redos, wrong, be.launches};
  memcpy(out, v, sizeof(v));
  return 0;
}
