// spec_dither_emul.cpp -- executable specification (CPU emulation) of the speculative segment-parallel
// Gilbert dither, SURVEY.md section 8(f) rank 1. STUDY INFRASTRUCTURE built on the oracle; not product code.
//
// The sequential dither (GC:187-280) carries three states from pixel to pixel:
//   Q  the error queue (ArrayDeque of DITHER_MAX boxes, GC:94)        -- re-synchronises (tools/resync_study.cpp)
//   D  how many java.util.Random.nextInt calls have been made (PL:467) -- a count, predictable per pixel
//   M  the first-seen memo nearestMap (PQ:271-274, PL:332-335,402)     -- grows, entries never change
// The curve is cut into segments. Every segment runs from a PREDICTED input state: Q from a warm-up of
// `warm` pixels started with an empty queue, D from a per-pixel prediction of the draws, M = the memo
// committed so far. Segments are then validated IN ORDER: segment s is exact iff its warmed-up Q equals
// the final Q of segment s-1 bit for bit, its D equals the D segment s-1 ended with, and every memo
// entry it created (a "claim") or read from the PREDICTED memo agrees with what the segments before it
// committed. The predicted memo is what a previous pass believed the first-seen entries to be, with the
// curve position of each (round 0: a pass with every error forced to zero; later: the claims of the last
// run of every segment); a lookup at position n may use a predicted entry whose position is < n.
// The first segment that fails is re-run in the next round from the exact state of its predecessor, later
// segments only if their D changed or a memo assumption collided. The result is bit-identical to the sequential
// run by construction; what this program measures is how many rounds and how much redundant work
// that takes, and it checks the assembled output against the sequential oracle.
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
namespace { struct ProbeLog { bool on = false; int n = 0; int32_t keys[8]; }; thread_local ProbeLog g_probe; }
#define NQ_ORACLE_PROBE(key) do { if (g_probe.on && g_probe.n < 8) g_probe.keys[g_probe.n++] = (key); } while (0)
#define private public
#define protected public
#define class struct
#include "../oracle/nq_oracle.cpp"
#undef private
#undef protected
#undef class
#include <unordered_set>

namespace {

struct EmuOut {
  int rounds = 0, nseg = 0, DM = 0, sorted = 0, exact = 0;
  long long pixelsRun = 0;        // pixels processed over all rounds, warm-up and re-runs included
  long long segRuns = 0;          // segment executions over all rounds
  long long claims = 0, draws = 0, specReads = 0;
  int failQ = 0, failD = 0, failM = 0;   // why validations failed
  int maxRerun = 0;
  long long errDependent = 0;     // pixels whose lookup colour differs from the zero-error pass (error-dependent lookups)
};
struct EmuCfg { int seg = 8192, warm = 2048, maxRounds = 4096; EmuOut out; };
EmuCfg* g_emu = nullptr;

using Box = GilbertCurve::ErrorBox;
struct Claim { int32_t key; short val; int pos; };

struct Seg {
  int p0 = 0, p1 = 0;
  bool hasRun = false, dirty = true, exactStart = false;
  std::deque<Box> Qstart;          // exact start state (when exactStart)
  long long Dbase = 0;             // draws before p0 assumed by the last run
  std::deque<Box> Qwarm, Qout;
  long long Dout = 0;
  std::vector<Claim> claims;       // memo entries created by owned pixels
  std::vector<Claim> reads;        // predicted memo entries read by owned pixels
  std::vector<int32_t> out;        // outputs of the owned pixels, curve order
};

uint64_t qhash(const std::deque<Box>& a) {
  uint64_t h = 1469598103934665603ULL;
  for (const Box& b : a) { uint32_t u[4]; memcpy(u, b.p, 16); for (int i = 0; i < 4; ++i) { h ^= u[i]; h *= 1099511628211ULL; } }
  return h;
}
bool same_queue(const std::deque<Box>& a, const std::deque<Box>& b) {
  if (a.size() != b.size()) return false;
  for (size_t i = 0; i < a.size(); ++i)
    if (memcmp(a[i].p, b[i].p, 16) != 0) return false;
  return true;
}

// Ditherable wrapper: sees which memo keys a lookup probed, created or read from the predicted set
struct Spy : Ditherable {
  Ditherable& inner;
  PnnQuantizer& q;
  std::vector<Claim>* reads = nullptr;     // reads of predicted entries
  std::unordered_set<int32_t>* spec = nullptr;
  int pos = 0;
  int32_t lastColour = 0;
  explicit Spy(Ditherable& in, PnnQuantizer& q_) : inner(in), q(q_) {}
  int getColorIndex(int32_t c) override { return inner.getColorIndex(c); }
  short nearestColorIndex(const std::vector<int32_t>& palette, int32_t c, int p) override {
    lastColour = c;
    const int i0 = g_probe.n;
    const short r = inner.nearestColorIndex(palette, c, p);
    if (spec && reads)
      for (int i = i0; i < g_probe.n; ++i) {
        const int32_t key = g_probe.keys[i];
        auto it = q.nearestMap.find(key);
        if (it != q.nearestMap.end() && spec->count(key)) reads->push_back(Claim{key, it->second, pos});
      }
    return r;
  }
};

std::vector<int32_t> emulate(PnnQuantizer& q, JRandom* rng, uint64_t rngSeed, int width, int height, const std::vector<int32_t>& cPixels,
                             std::vector<int32_t>& palette, Ditherable& inner, const std::vector<float>* sal, double weight, bool dither,
                             const std::vector<int32_t>& reference) {
  EmuCfg& cfg = *g_emu;
  EmuOut& R = cfg.out;
  const int npix = width * height;
  OrderOnly oo;
  oo.width = width;
  oo.out.reserve(npix);
  if (width >= height) oo.gen(0, 0, width, 0, 0, height); else oo.gen(0, 0, 0, height, width, 0);
  const std::vector<uint32_t>& order = oo.out;
  Spy spy(inner, q);
  g_probe.on = true;
  std::vector<int32_t> scratch(npix, 0);

  {
    GilbertCurve probe(width, height, cPixels, palette, scratch, spy, sal, weight, dither);
    R.DM = probe.DITHER_MAX; R.sorted = probe.sortedByYDiff;
    if (probe.sortedByYDiff) return {};   // PriorityQueue mode is not covered by this scheme
  }

  // java.util.Random state after j nextInt(32767) calls ("the j-th call returns the j-th non-rejected value")
  std::vector<uint64_t> rngState;
  if (rng) {
    rngState.resize((size_t)npix * 3 + 16);
    JRandom r(0);
    r.setSeed(rngSeed);
    for (size_t j = 0; j < rngState.size(); ++j) { rngState[j] = r.seed; r.nextInt(32767); }
  }

  // ---- round 0: zero-error pass (all weights 0: every lookup sees the undiffused pixel). Gives the per-pixel draw
  //      prediction (PL:467 draws iff closest[2] != 0, whatever the value drawn), the predicted memo with positions,
  //      and the lookup colour of every pixel (to count error-dependent lookups afterwards).
  std::vector<uint8_t> drawn(npix, 0);
  std::vector<int32_t> zeroColour(npix, 0);
  std::unordered_map<int32_t, Claim> Mpred;
  {
    GilbertCurve g0(width, height, cPixels, palette, scratch, spy, sal, weight, dither);
    g0.initWeights(g0.DITHER_MAX);
    for (auto& w : g0.weights) w = 0.f;
    if (rng) rng->seed = rngState[0];
    for (int n = 0; n < npix; ++n) {
      const int bidx = (int)order[n];
      scratch[bidx] = 0;
      const uint64_t s0 = rng ? rng->seed : 0;
      const size_t m0 = q.nearestMap.size();
      spy.pos = n;
      g_probe.n = 0;
      g0.diffusePixel(bidx % width, bidx / width);
      drawn[n] = (rng && rng->seed != s0) ? 1 : 0;
      zeroColour[n] = spy.lastColour;
      if (q.nearestMap.size() > m0)
        for (int i = 0; i < g_probe.n; ++i) {
          const int32_t key = g_probe.keys[i];
          auto it = q.nearestMap.find(key);
          if (it != q.nearestMap.end() && !Mpred.count(key)) Mpred.emplace(key, Claim{key, it->second, n});
        }
    }
    q.nearestMap.clear();
  }

  const int nseg = (npix + cfg.seg - 1) / cfg.seg;
  R.nseg = nseg;
  std::vector<Seg> S(nseg);
  for (int s = 0; s < nseg; ++s) { S[s].p0 = s * cfg.seg; S[s].p1 = std::min(npix, (s + 1) * cfg.seg); }
  S[0].exactStart = true;   // the first segment starts from the true initial state (DITHER_MAX empty boxes, GC:358-359)

  std::vector<int32_t> trueColour(npix, 0);
  std::vector<uint64_t> trueHash;
  bool recordHash = false;

  auto run_segment = [&](Seg& sg, long long Dbase) {
    GilbertCurve g(width, height, cPixels, palette, scratch, spy, sal, weight, dither);
    g.initWeights(g.DITHER_MAX);            // weights + DITHER_MAX empty boxes
    int from = sg.p0;
    if (sg.exactStart) { if (sg.p0 > 0) g.fifo = sg.Qstart; }
    else from = std::max(0, sg.p0 - cfg.warm);
    long long D = Dbase;
    for (int n = from; n < sg.p0; ++n) D -= drawn[n];
    if (D < 0) D = 0;
    // predicted entries by position, those not committed yet
    std::vector<const Claim*> pred;
    // only entries first seen BEFORE the segment: what lies inside it, the segment creates itself when it gets there
    for (auto& kv : Mpred) if (kv.second.pos < sg.p0 && !q.nearestMap.count(kv.first)) pred.push_back(&kv.second);
    std::sort(pred.begin(), pred.end(), [](const Claim* a, const Claim* b) { return a->pos < b->pos; });
    size_t pi = 0;
    std::unordered_set<int32_t> spec;
    std::vector<int32_t> added;             // every key this run put into q.nearestMap (restored at the end)
    sg.claims.clear(); sg.reads.clear();
    sg.out.assign(sg.p1 - sg.p0, 0);
    std::vector<uint8_t> myDrawn(sg.p1 - sg.p0, 0);
    spy.spec = &spec; spy.reads = nullptr;
    if (rng) rng->seed = rngState[D];
    for (int n = from; n < sg.p1; ++n) {
      if (n == sg.p0) {
        sg.Qwarm = g.fifo;
        // warm-up knowledge is local and dropped: entries it created itself disappear
        for (int32_t k : added) if (!spec.count(k)) q.nearestMap.erase(k);
        std::vector<int32_t> keep;
        for (int32_t k : added) if (spec.count(k)) keep.push_back(k);
        added.swap(keep);
        for (size_t k = 0; k < pi; ++k)      // predicted entries the warm-up created itself (their position lies inside it)
          if (q.nearestMap.emplace(pred[k]->key, pred[k]->val).second) { spec.insert(pred[k]->key); added.push_back(pred[k]->key); }
        spy.reads = &sg.reads;
        D = Dbase;
        if (rng) rng->seed = rngState[D];
      }
      if (n == from)               // every predicted entry first seen before the segment, the warm-up window included: the
        while (pi < pred.size()) { // warm-up wants the best guess, not first-seen order
          const Claim* c = pred[pi++];
          if (q.nearestMap.emplace(c->key, c->val).second) { spec.insert(c->key); added.push_back(c->key); }
        }
      const int bidx = (int)order[n];
      scratch[bidx] = 0;
      spy.pos = n;
      const uint64_t s0 = rng ? rng->seed : 0;
      const size_t m0 = q.nearestMap.size();
      g_probe.n = 0;
      g.diffusePixel(bidx % width, bidx / width);
      if (q.nearestMap.size() > m0)
        for (int i = 0; i < g_probe.n; ++i) {
          const int32_t key = g_probe.keys[i];
          auto it = q.nearestMap.find(key);
          if (it == q.nearestMap.end() || spec.count(key)) continue;
          bool mine = false;
          for (int32_t k : added) if (k == key) { mine = true; break; }
          if (mine) continue;
          added.push_back(key);
          if (n >= sg.p0) sg.claims.push_back(Claim{key, it->second, n});
        }
      if (recordHash) trueHash[n] = qhash(g.fifo);
      else if (!trueHash.empty() && getenv("NQ_EMU_DEBUG") && n < sg.p0 && ((n - from) % 512 == 511 || n == sg.p0 - 1))
        fprintf(stderr, "seg %d warm-up pixel %d (+%d): state %s, lookup colour %s\n", sg.p0 / cfg.seg, n, n - from, qhash(g.fifo) == trueHash[n] ? "SAME" : "differs", spy.lastColour == trueColour[n] ? "same" : "differs");
      if (rng && rng->seed != s0) { ++D; if (n >= sg.p0) myDrawn[n - sg.p0] = 1; }
      if (n >= sg.p0) { sg.out[n - sg.p0] = scratch[bidx]; trueColour[n] = spy.lastColour; }
    }
    sg.Qout = g.fifo;
    sg.Dbase = Dbase;
    sg.Dout = D;
    for (int32_t k : added) q.nearestMap.erase(k);       // back to the committed memo
    for (int n = sg.p0; n < sg.p1; ++n) drawn[n] = myDrawn[n - sg.p0];
    spy.spec = nullptr; spy.reads = nullptr;
    sg.hasRun = true; sg.dirty = false;
    R.pixelsRun += sg.p1 - from;
    ++R.segRuns;
  };

  std::unordered_map<int32_t, int> commitPos;
  std::unordered_map<int32_t, Claim> Mzero = Mpred;
  auto memo_ok = [&](const Seg& sg) {
    for (const Claim& c : sg.reads) {           // a predicted entry it used must have been committed with that value
      auto it = q.nearestMap.find(c.key);
      if (it == q.nearestMap.end() || it->second != c.val) {
        if (getenv("NQ_EMU_DEBUG")) fprintf(stderr, "seg %d: read key %x val %d pos %d: committed %s %d\n", sg.p0 / g_emu->seg, c.key, c.val, c.pos, it == q.nearestMap.end() ? "absent" : "has", it == q.nearestMap.end() ? -1 : it->second);
        return false;
      }
    }
    for (const Claim& c : sg.claims) {          // an entry it created must not exist with another value
      auto it = q.nearestMap.find(c.key);
      if (it != q.nearestMap.end() && it->second != c.val) {
        if (getenv("NQ_EMU_DEBUG")) fprintf(stderr, "seg %d: claim key %x val %d pos %d: committed %d at pos %d; zero-pass predicted: %s pos %d val %d\n", sg.p0 / g_emu->seg, c.key, c.val, c.pos, it->second, commitPos[c.key],
            Mzero.count(c.key) ? "yes" : "no", Mzero.count(c.key) ? Mzero[c.key].pos : -1, Mzero.count(c.key) ? Mzero[c.key].val : -1);
        return false;
      }
    }
    return true;
  };

  if (getenv("NQ_EMU_ORACLE_MEMO")) {   // experiment: predict the memo perfectly (claims of one exact run of the whole curve)
    Seg all; all.p0 = 0; all.p1 = npix; all.exactStart = true;
    Mpred.clear();
    trueHash.assign(npix, 0); recordHash = true;
    run_segment(all, 0);
    recordHash = false;
    for (const Claim& c : all.claims) Mpred.emplace(c.key, c);
    R.pixelsRun = 0; R.segRuns = 0;
    if (rng) { /* keep the zero-pass draw prediction */ }
  }
  int firstOpen = 0;             // segments [0, firstOpen) are validated and committed
  std::deque<Box> Qprev;         // final state of segment firstOpen-1
  long long Dprev = 0;
  std::vector<int32_t> result(npix, 0);
  while (firstOpen < nseg && R.rounds < cfg.maxRounds) {
    ++R.rounds;
    // ---- "parallel" phase: every dirty segment runs against the same committed + predicted memo, from a draw base
    //      predicted with what is known BEFORE the round (last run's count, else the zero-error prediction)
    std::vector<long long> Dpred(nseg, 0);
    {
      long long d = Dprev;
      for (int s = firstOpen; s < nseg; ++s) {
        Dpred[s] = d;
        if (S[s].hasRun) d += S[s].Dout - S[s].Dbase;
        else for (int n = S[s].p0; n < S[s].p1; ++n) d += drawn[n];
      }
    }
    int rerun = 0;
    for (int s = firstOpen; s < nseg; ++s) {
      Seg& sg = S[s];
      if (sg.hasRun && !sg.dirty && sg.Dbase != Dpred[s]) sg.dirty = true;   // its draw base moved
      if (sg.dirty) { run_segment(sg, Dpred[s]); ++rerun; }
    }
    R.maxRerun = std::max(R.maxRerun, rerun);
    // ---- ordered validation
    for (int s = firstOpen; s < nseg; ++s) {
      Seg& sg = S[s];
      bool ok = true;
      if (s > 0) {
        if (!same_queue(sg.exactStart ? sg.Qstart : sg.Qwarm, Qprev)) {
          ok = false; ++R.failQ;
          if (getenv("NQ_EMU_DEBUG")) fprintf(stderr, "round %d: seg %d fails on Q (exactStart %d, hash warm %llx prev %llx true %llx)\n", R.rounds, s, (int)sg.exactStart,
              (unsigned long long)qhash(sg.exactStart ? sg.Qstart : sg.Qwarm), (unsigned long long)qhash(Qprev), trueHash.empty() ? 0ULL : (unsigned long long)trueHash[sg.p0 - 1]);
        }
        if (ok && sg.Dbase != Dprev) { ok = false; ++R.failD; }
      }
      if (ok && !memo_ok(sg)) { ok = false; ++R.failM; }
      if (!ok) {
        sg.exactStart = true; sg.Qstart = Qprev; sg.dirty = true;
        break;
      }
      for (const Claim& c : sg.claims) { if (q.nearestMap.emplace(c.key, c.val).second) commitPos[c.key] = c.pos; ++R.claims; }
      R.specReads += (long long)sg.reads.size();
      for (int n = sg.p0; n < sg.p1; ++n) result[order[n]] = sg.out[n - sg.p0];
      Qprev = sg.Qout; Dprev = sg.Dout;
      firstOpen = s + 1;
    }
    // ---- new prediction: the claims of the last run of every open segment, first position wins
    Mpred.clear();
    for (int s = firstOpen; s < nseg; ++s)
      for (const Claim& c : S[s].claims) if (!q.nearestMap.count(c.key) && !Mpred.count(c.key)) Mpred.emplace(c.key, c);
    // segments whose memo assumptions no longer hold against committed + newly predicted entries re-run
    for (int s = firstOpen; s < nseg; ++s) {
      if (S[s].dirty) continue;
      for (const Claim& c : S[s].reads) {
        auto it = q.nearestMap.find(c.key);
        if (it != q.nearestMap.end()) { if (it->second != c.val) { S[s].dirty = true; break; } continue; }
        auto ip = Mpred.find(c.key);
        if (ip == Mpred.end() || ip->second.val != c.val || ip->second.pos >= c.pos) { S[s].dirty = true; break; }
      }
      if (S[s].dirty) continue;
      for (const Claim& c : S[s].claims) {
        auto it = q.nearestMap.find(c.key);
        if (it != q.nearestMap.end() && it->second != c.val) { S[s].dirty = true; break; }
        auto ip = Mpred.find(c.key);
        if (ip != Mpred.end() && ip->second.pos < c.pos && ip->second.val != c.val) { S[s].dirty = true; break; }
      }
    }
  }
  R.draws = Dprev;
  R.exact = firstOpen == nseg && result == reference;
  if (R.exact) for (int n = 0; n < npix; ++n) R.errDependent += trueColour[n] != zeroColour[n];
  return result;
}

struct EmuRgb : PnnQuantizer {
  using PnnQuantizer::PnnQuantizer;
  std::vector<int32_t> ditherImage(const std::vector<int32_t>& cPixels, std::vector<int32_t>& palette, int width, int height, bool dither) override {
    const double w0 = weight;
    std::vector<int32_t> ref = PnnQuantizer::ditherImage(cPixels, palette, width, height, dither);
    weight = w0;
    RgbDitherable ditherable(*this, dither);
    if (hasSemiTransparency) weight *= -1;
    if (!dither && palette.size() > 32) {
      // dither == false: only the Gilbert pass is emulated (the BlueNoise second pass, BN:207-222, is a raster pass of
      // its own); its sequential result = one GilbertCurve run with fresh maps
      std::vector<int32_t> ref1(cPixels.size(), 0);
      { GilbertCurve gc(width, height, cPixels, palette, ref1, ditherable, nullptr, weight, dither); gc.run(); }
      closestMap.clear(); nearestMap.clear();
      emulate(*this, nullptr, 0, width, height, cPixels, palette, ditherable, nullptr, weight, dither, ref1);
      return ref;
    }
    emulate(*this, nullptr, 0, width, height, cPixels, palette, ditherable, nullptr, weight, dither, ref);
    return ref;
  }
};
struct EmuLab : PnnLABQuantizer {
  using PnnLABQuantizer::PnnLABQuantizer;
  std::vector<int32_t> ditherImage(const std::vector<int32_t>& cPixels, std::vector<int32_t>& palette, int width, int height, bool dither) override {
    const double w0 = weight;
    std::vector<int32_t> ref = PnnLABQuantizer::ditherImage(cPixels, palette, width, height, dither);
    weight = w0;
    if (!dither && palette.size() > 32) return ref;
    LabDitherable ditherable(*this);
    if (hasSemiTransparency) weight *= -1;
    emulate(*this, &random, rngSeed, width, height, cPixels, palette, ditherable, hasSaliencies ? &saliencies : nullptr, weight, dither, ref);
    return ref;
  }
};

}  // namespace

extern "C" int nqs_spec_emulate(int kind, const uint32_t* argb, int w, int h, int nmax, int dither, uint64_t seed, int seg, int warm,
                                long long* out /* 16 values */) {
  EmuCfg cfg;
  cfg.seg = seg; cfg.warm = warm;
  g_emu = &cfg;
  M.mode = 0;
  try {
    if (kind == 0) { EmuRgb q(argb, w, h); q.rngSeed = seed; q.convert(nmax, dither != 0); }
    else { EmuLab q(argb, w, h); q.rngSeed = seed; q.convert(nmax, dither != 0); }
  } catch (const std::exception& e) {
    fprintf(stderr, "emulate: %s\n", e.what());
    g_emu = nullptr;
    return -1;
  }
  const EmuOut& R = cfg.out;
  long long v[16] = {R.rounds, R.nseg, R.DM, R.sorted, R.exact, R.pixelsRun, R.segRuns, R.claims, R.draws, R.failQ, R.failD, R.failM, R.maxRerun, R.specReads, R.errDependent, 0};
  memcpy(out, v, sizeof(v));
  g_emu = nullptr;
  return 0;
}
