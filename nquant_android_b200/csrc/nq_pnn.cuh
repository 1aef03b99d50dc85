// nq_pnn.cuh -- pairwise-nearest-neighbour clustering of the compacted bins.
// Reference: find_nn (PQ:57-116, PL:44-115), heap build (PQ:195-207, PL:245-257), merge loop
// (PQ:210-255, PL:267-312), palette fill (PQ:258-264, PL:315-324).
//
// What must be preserved for the merge sequence (and therefore the palette) to be identical:
//  * find_nn is an ORDERED scan over the surviving bins after idx: whether candidate i is taken
//    depends on the running err left by earlier candidates. RGB: a candidate that passes the four
//    `continue` tests is always taken, with err := the first partial sum of the YUV loop that
//    reaches err (the `break` falls through, PQ:97-112) -- err can grow. LAB: err := full cost,
//    early exits use >= for the first tests and > later, and R_T may be negative (PL:57-108).
//    We evaluate the err-independent partial sums of many candidates in parallel and resolve
//    acceptance in list order with ballot rounds (one round per ACCEPTED candidate).
//  * the binary heap code, including its tie behaviour, is replayed verbatim by one thread; heap
//    slots carry a copy of the bin's err (a bin's err only changes while it sits at heap[1]).
//  * forward links only ever skip deleted bins, so "list order" == ascending index among live bins;
//    a periodically compacted live list replaces pointer chasing.
#pragma once
#include "nq_types.h"
#include "nq_color.h"
#include "nq_hist.cuh"

namespace nq {

__device__ const float c_coeffs[3][3] = {{0.299f, 0.587f, 0.114f}, {-0.14713f, -0.28886f, 0.436f}, {0.615f, -0.51499f, -0.10001f}};  // PQ:26-30

// -------------------------------------------------------------------------------------------------
// RGB candidate arithmetic (PQ:73-112)
// -------------------------------------------------------------------------------------------------
struct RgbProbe {
  float n1;
  double wa, wr, wg, wb;
  double PR, PG, PB, PA, ratio;
  int start;
  bool semi;
};

__device__ __forceinline__ RgbProbe rgb_probe(const NqImage& I, const NqSlot& S, int idx) {
  RgbProbe P;
  P.n1 = S.bCnt[idx];
  P.wa = S.bAc[idx]; P.wr = S.bC1[idx]; P.wg = S.bC2[idx]; P.wb = S.bC3[idx];
  P.PR = I.PR; P.PG = I.PG; P.PB = I.PB; P.PA = I.PA; P.ratio = I.ratio;
  P.semi = I.hasSemi != 0;
  P.start = 0;
  if (g_blueNoise[idx & 4095] > 0) P.start = (I.PG < (double)0.587f) ? 3 : 1;   // PQ:69-71
  return P;
}

struct RgbCand { double nerr2, q0, dr, dg, db; };

// gate: candidate is taken iff nerr2 < err and q0 < err (the partial sums before q0 are <= q0)
__device__ __forceinline__ double rgb_gate(const RgbProbe& P, const NqSlot& S, int i, RgbCand* c) {
  double n2 = (double)S.bCnt[i];
  double nerr2 = ((double)P.n1 * n2) / ((double)P.n1 + n2);
  double nerr = 0.0;
  if (P.semi) { double d = S.bAc[i] - P.wa; nerr += nerr2 * P.PA * (d * d); }
  double dr = S.bC1[i] - P.wr, dg = S.bC2[i] - P.wg, db = S.bC3[i] - P.wb;
  nerr += nerr2 * (1 - P.ratio) * P.PR * (dr * dr);
  nerr += nerr2 * (1 - P.ratio) * P.PG * (dg * dg);
  nerr += nerr2 * (1 - P.ratio) * P.PB * (db * db);
  c->nerr2 = nerr2; c->q0 = nerr; c->dr = dr; c->dg = dg; c->db = db;
  return nerr2 > nerr ? nerr2 : nerr;
}
// new err of a taken candidate (PQ:97-112)
__device__ __forceinline__ double rgb_take(const RgbProbe& P, const RgbCand& c, double err) {
  double nerr = c.q0;
  for (int j = P.start; j < 3; ++j) {
    double t = (double)c_coeffs[j][0] * c.dr;
    nerr += c.nerr2 * P.ratio * (t * t);
    if (nerr >= err) break;
    t = (double)c_coeffs[j][1] * c.dg;
    nerr += c.nerr2 * P.ratio * (t * t);
    if (nerr >= err) break;
    t = (double)c_coeffs[j][2] * c.db;
    nerr += c.nerr2 * P.ratio * (t * t);
    if (nerr >= err) break;
  }
  return nerr;
}

// -------------------------------------------------------------------------------------------------
// LAB candidate arithmetic (PL:54-112). err only ever decreases, so a candidate that fails a test
// against the err at the start of a chunk can be dropped for good. A survivor carries
// gs (must be < err), gw (must be <= err) and f (the new err).
// -------------------------------------------------------------------------------------------------
struct LabProbe {
  float n1, a1, L1, A1, B1;
  double ratio, exp175;
  bool semi, texicab;
};
__device__ __forceinline__ LabProbe lab_probe(const NqImage& I, const NqSlot& S, int idx, double ratio) {
  LabProbe P;
  P.n1 = S.bCnt[idx];
  P.a1 = S.fAc[idx]; P.L1 = S.fC1[idx]; P.A1 = S.fC2[idx]; P.B1 = S.fC3[idx];
  P.ratio = ratio;
  P.semi = I.hasSemi != 0;
  P.exp175 = P.semi ? nqm::nq_exp(1.75) : 1.0;
  P.texicab = I.texicab != 0;
  return P;
}
struct LabCand { double gs, gw, f; };

__device__ __forceinline__ bool lab_eval(const LabProbe& P, const NqSlot& S, int i, double err, LabCand* c) {
  float n2 = S.bCnt[i];
  double nerr2 = (double)((P.n1 * n2) / (P.n1 + n2));
  if (nerr2 >= err) return false;
  float a2 = S.fAc[i], L2 = S.fC1[i], A2 = S.fC2[i], B2 = S.fC3[i];
  double alphaDiff = 0;
  if (P.semi) { double d = (double)(a2 - P.a1); alphaDiff = (d * d) / P.exp175; }
  double nerr = nerr2 * alphaDiff;
  if (nerr >= err) return false;
  double gs = nerr2;
  if (!P.texicab) {
    double d = (double)(L2 - P.L1);
    nerr += (1 - P.ratio) * nerr2 * (d * d);
    if (nerr >= err) return false;
    d = (double)(A2 - P.A1);
    nerr += (1 - P.ratio) * nerr2 * (d * d);
    if (nerr >= err) return false;
    gs = nerr > gs ? nerr : gs;
    d = (double)(B2 - P.B1);
    nerr += (1 - P.ratio) * nerr2 * (d * d);
  } else {
    nerr += (1 - P.ratio) * nerr2 * (double)fabsf(L2 - P.L1);
    if (nerr >= err) return false;
    gs = nerr > gs ? nerr : gs;
    double da = (double)(A2 - P.A1), db = (double)(B2 - P.B1);
    nerr += (1 - P.ratio) * nerr2 * nqm::sqrt_((da * da) + (db * db));
  }
  if (nerr > err) return false;

  float tL = ciede_L(P.L1, L2);
  nerr += P.ratio * nerr2 * ((double)tL * (double)tL);
  if (nerr > err) return false;

  CiedeC cc;
  float tC = ciede_C(P.A1, P.B1, A2, B2, &cc);
  nerr += P.ratio * nerr2 * ((double)tC * (double)tC);
  if (nerr > err) return false;

  double barC, barh;
  float tH = ciede_H(P.B1, B2, cc, &barC, &barh);
  nerr += P.ratio * nerr2 * ((double)tH * (double)tH);
  if (nerr > err) return false;
  double gw = nerr;

  nerr += P.ratio * nerr2 * (double)ciede_RT(barC, barh, tC, tH);
  if (nerr > err) return false;
  c->gs = gs; c->gw = nerr > gw ? nerr : gw; c->f = nerr;
  return true;
}

// -------------------------------------------------------------------------------------------------
// initial sweep: find_nn for every bin (PQ:196-197, PL:246-247). One warp per bin.
// -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_find_nn_all(NqImage* imgs, const NqSlot* slots, int nimg) {
  const unsigned lane = lane_id();
  const int wpb = blockDim.x >> 5;
  for (int img = 0; img < nimg; ++img) {
    NqImage& I = imgs[img];
    if (I.nmax <= 2 || I.skipPnn) continue;
    const NqSlot& S = slots[img];
    const int maxbins = I.maxbins;
    unsigned long long pairs = 0;
    for (int idx = blockIdx.x * wpb + (threadIdx.x >> 5); idx < maxbins; idx += gridDim.x * wpb) {
      double err = 1e100;
      int nn = 0;
      if (I.kind == NQ_KIND_RGB) {
        RgbProbe P = rgb_probe(I, S, idx);
        for (int base = idx + 1; base < maxbins; base += 32) {
          const int i = base + lane;
          RgbCand c;
          double gate = 1e300;
          if (i < maxbins) gate = rgb_gate(P, S, i, &c);
          unsigned remaining = 0xffffffffu;
          for (;;) {
            unsigned m = __ballot_sync(0xffffffffu, gate < err) & remaining;
            if (!m) break;
            int L = __ffs(m) - 1;
            double e = 0;
            if ((int)lane == L) e = rgb_take(P, c, err);
            err = __shfl_sync(0xffffffffu, e, L);
            nn = base + L;
            remaining = (L == 31) ? 0u : ~((2u << L) - 1u);
          }
        }
      } else {
        LabProbe P = lab_probe(I, S, idx, I.ratio);
        for (int base = idx + 1; base < maxbins; base += 32) {
          const int i = base + lane;
          LabCand c;
          bool alive = false;
          if (i < maxbins) alive = lab_eval(P, S, i, err, &c);
          unsigned remaining = 0xffffffffu;
          for (;;) {
            unsigned m = __ballot_sync(0xffffffffu, alive && c.gs < err && c.gw <= err) & remaining;
            if (!m) break;
            int L = __ffs(m) - 1;
            err = __shfl_sync(0xffffffffu, c.f, L);
            nn = base + L;
            remaining = (L == 31) ? 0u : ~((2u << L) - 1u);
          }
        }
      }
      pairs += (unsigned long long)(maxbins - idx - 1);
      if (lane == 0) { S.bErr[idx] = (float)err; S.bNn[idx] = nn; }
    }
    if (lane == 0 && pairs) atomicAdd(&I.statPairs, pairs);
  }
}

// -------------------------------------------------------------------------------------------------
// merge loop: one persistent CTA per image.
// -------------------------------------------------------------------------------------------------
#define NQ_MERGE_THREADS 1024
#define NQ_HEAP_SMEM 32768     // heap slots kept in shared memory (err f32 + id u16 = 6 B each = 192 KB)

struct HeapView {
  float* sErr; unsigned short* sId;   // shared part, slots [0, NQ_HEAP_SMEM)
  float* gErr; int* gId;              // global spill for the deepest level(s)
  __device__ __forceinline__ float err(int l) const { return l < NQ_HEAP_SMEM ? sErr[l] : gErr[l]; }
  __device__ __forceinline__ int id(int l) const { return l < NQ_HEAP_SMEM ? (int)sId[l] : gId[l]; }
  __device__ __forceinline__ void set(int l, int id_, float e) {
    if (l < NQ_HEAP_SMEM) { sErr[l] = e; sId[l] = (unsigned short)id_; } else { gErr[l] = e; gId[l] = id_; }
  }
  // "push slot down" (PQ:228-236): sift (b1, e1) down from the root of a heap with heapN entries
  __device__ __forceinline__ void sift_down(int b1, float e1, int heapN) {
    int l = 1, l2;
    for (; (l2 = l + l) <= heapN; l = l2) {
      float ea = err(l2);
      if (l2 < heapN) { float eb = err(l2 + 1); if (ea > eb) { ++l2; ea = eb; } }
      if (e1 <= ea) break;
      set(l, id(l2), ea);
    }
    set(l, b1, e1);
  }
};

// rebuild the ascending list of live bins; returns its length (block-wide, all threads call)
__device__ __forceinline__ int rebuild_live(const NqSlot& S, int maxbins, int* live, int* posOf, int* sWarp) {
  const int t = threadIdx.x;
  const int per = (maxbins + NQ_MERGE_THREADS - 1) / NQ_MERGE_THREADS;
  const int b0 = min(maxbins, t * per), b1 = min(maxbins, b0 + per);
  int c = 0;
  for (int b = b0; b < b1; ++b) c += S.bMtm[b] != NQ_DELETED;
  int total, j = block_excl_scan_1024(c, &total, sWarp);
  for (int b = b0; b < b1; ++b)
    if (S.bMtm[b] != NQ_DELETED) { live[j] = b; posOf[b] = j; ++j; }
  __syncthreads();
  return total;
}

__global__ void __launch_bounds__(NQ_MERGE_THREADS, 1) k_merge(NqImage* imgs, const NqSlot* slots, int* liveBuf, int* posBuf, int logMerges) {
  extern __shared__ unsigned char smemRaw[];
  float* sErr = reinterpret_cast<float*>(smemRaw);
  unsigned short* sId = reinterpret_cast<unsigned short*>(smemRaw + (size_t)NQ_HEAP_SMEM * 4);
  __shared__ int sWarp[33];
  __shared__ unsigned sMask[32];
  __shared__ double sErrCur;
  __shared__ int sNnCur, sAction, sB1, sHeapN, sIter;

  const int img = blockIdx.x;
  NqImage& I = imgs[img];
  if (I.nmax <= 2 || I.skipPnn) return;
  const NqSlot& S = slots[img];
  const int t = threadIdx.x;
  const unsigned lane = lane_id(), w = t >> 5;
  const int maxbins = I.maxbins, extbins = I.extbins;
  const bool rgb = I.kind == NQ_KIND_RGB;
  int* live = liveBuf + (size_t)img * NQ_NBINS;
  int* posOf = posBuf + (size_t)img * NQ_NBINS;
  HeapView H{sErr, sId, S.hErr, S.hId};

  // ---- heap build: sequential pushes in bin order (PQ:196-207). Warp 0 replays them; every lane
  //      follows the same scalar steps (values prefetched 32 at a time), lane 0 does the stores.
  if (w == 0) {
    int heapN = 0;
    for (int base = 0; base < maxbins; base += 32) {
      float e = (base + (int)lane < maxbins) ? S.bErr[base + lane] : 0.f;
      const int cnt = min(32, maxbins - base);
      for (int j = 0; j < cnt; ++j) {
        const float err = __shfl_sync(0xffffffffu, e, j);
        int l = ++heapN, l2;
        for (; l > 1; l = l2) {
          l2 = l >> 1;
          float pe = H.err(l2);
          if (pe <= err) break;
          int pid = H.id(l2);
          if (lane == 0) H.set(l, pid, pe);
        }
        if (lane == 0) H.set(l, base + j, err);
        __syncwarp();
      }
    }
    if (lane == 0) { sHeapN = heapN; sIter = 0; }
  }
  __syncthreads();
  int liveLen = rebuild_live(S, maxbins, live, posOf, sWarp);
  int liveAtRebuild = liveLen, iterAtRebuild = 0;
  unsigned long long rescans = 0, pairs = 0;
  unsigned pops = 0;
  const double ratioMerge = I.ratioMerge;

  for (;;) {
    // ---- thread 0: look at the heap top (PQ:214-226)
    if (t == 0) {
      int action = 0;  // 0 = merge, 1 = rescan, 2 = finished
      if (sIter >= extbins) action = 2;
      else {
        int heapN = sHeapN;
        for (;;) {
          int b1 = H.id(1);
          int tm = S.bTm[b1], mtm = S.bMtm[b1];
          if (tm >= mtm && S.bMtm[S.bNn[b1]] <= tm) { action = 0; sB1 = b1; break; }
          if (mtm == NQ_DELETED) {   // deleted node: b1 = heap[1] = heap[heap[0]--], then push down
            b1 = H.id(heapN);
            float e1 = H.err(heapN);
            --heapN;
            ++pops;
            H.sift_down(b1, e1, heapN);
            continue;
          }
          action = 1; sB1 = b1;
          break;
        }
        sHeapN = heapN;
      }
      sAction = action;
    }
    __syncthreads();
    const int action = sAction;
    if (action == 2) break;
    const int b1 = sB1;

    if (action == 1) {
      // ---- cooperative find_nn(b1) over the live bins after b1
      ++rescans;
      const int p0 = posOf[b1] + 1;
      double err = 1e100;
      int nn = 0;
      if (rgb) {
        RgbProbe P = rgb_probe(I, S, b1);
        for (int base = p0; base < liveLen; base += NQ_MERGE_THREADS) {
          const int p = base + t;
          int i = -1;
          RgbCand c;
          double gate = 1e300;
          if (p < liveLen) {
            i = live[p];
            if (S.bMtm[i] != NQ_DELETED) gate = rgb_gate(P, S, i, &c); else i = -1;
          }
          int lastP = -1;
          for (;;) {
            unsigned m = __ballot_sync(0xffffffffu, i >= 0 && p > lastP && gate < err);
            if (lane == 0) sMask[w] = m;
            __syncthreads();
            unsigned mine = sMask[lane];
            unsigned wm = __ballot_sync(0xffffffffu, mine != 0);
            if (!wm) { __syncthreads(); break; }
            int fw = __ffs(wm) - 1;
            unsigned mm = __shfl_sync(0xffffffffu, mine, fw);
            int winner = fw * 32 + (__ffs(mm) - 1);
            if (t == winner) { sErrCur = rgb_take(P, c, err); sNnCur = i; }
            __syncthreads();
            err = sErrCur; nn = sNnCur;
            lastP = base + winner;
          }
        }
      } else {
        LabProbe P = lab_probe(I, S, b1, ratioMerge);
        for (int base = p0; base < liveLen; base += NQ_MERGE_THREADS) {
          const int p = base + t;
          int i = -1;
          LabCand c;
          bool ok = false;
          if (p < liveLen) {
            i = live[p];
            if (S.bMtm[i] != NQ_DELETED) ok = lab_eval(P, S, i, err, &c);
          }
          int lastP = -1;
          for (;;) {
            unsigned m = __ballot_sync(0xffffffffu, ok && p > lastP && c.gs < err && c.gw <= err);
            if (lane == 0) sMask[w] = m;
            __syncthreads();
            unsigned mine = sMask[lane];
            unsigned wm = __ballot_sync(0xffffffffu, mine != 0);
            if (!wm) { __syncthreads(); break; }
            int fw = __ffs(wm) - 1;
            unsigned mm = __shfl_sync(0xffffffffu, mine, fw);
            int winner = fw * 32 + (__ffs(mm) - 1);
            if (t == winner) { sErrCur = c.f; sNnCur = i; }
            __syncthreads();
            err = sErrCur; nn = sNnCur;
            lastP = base + winner;
          }
        }
      }
      pairs += (unsigned long long)max(0, liveLen - p0);
      if (t == 0) {
        // tb.tm = i; push slot down (PQ:224-236)
        float e1 = (float)err;
        S.bErr[b1] = e1; S.bNn[b1] = nn; S.bTm[b1] = sIter;
        H.sift_down(b1, e1, sHeapN);
      }
      __syncthreads();
      continue;
    }

    // ---- merge tb <- tb + nb (PQ:240-254, PL:297-311)
    if (t == 0) {
      const int nbI = S.bNn[b1];
      const float n1 = S.bCnt[b1], n2 = S.bCnt[nbI];
      if (rgb) {
        const float d = 1.f / (n1 + n2);
        S.bAc[b1] = (double)(d * (float)jround((double)n1 * S.bAc[b1] + (double)n2 * S.bAc[nbI]));   // float * long -> float
        S.bC1[b1] = (double)(d * (float)jround((double)n1 * S.bC1[b1] + (double)n2 * S.bC1[nbI]));   // float * long -> float
        S.bC2[b1] = (double)(d * (float)jround((double)n1 * S.bC2[b1] + (double)n2 * S.bC2[nbI]));   // float * long -> float
        S.bC3[b1] = (double)(d * (float)jround((double)n1 * S.bC3[b1] + (double)n2 * S.bC3[nbI]));   // float * long -> float
      } else {
        const float d = 1.0f / (n1 + n2);
        S.fAc[b1] = d * (n1 * S.fAc[b1] + n2 * S.fAc[nbI]);
        S.fC1[b1] = d * (n1 * S.fC1[b1] + n2 * S.fC1[nbI]);
        S.fC2[b1] = d * (n1 * S.fC2[b1] + n2 * S.fC2[nbI]);
        S.fC3[b1] = d * (n1 * S.fC3[b1] + n2 * S.fC3[nbI]);
      }
      S.bCnt[b1] = n1 + n2;
      const int i = sIter + 1;
      S.bMtm[b1] = i;
      S.bMtm[nbI] = NQ_DELETED;
      if (logMerges && S.mergeLog) { S.mergeLog[2 * (i - 1)] = b1; S.mergeLog[2 * (i - 1) + 1] = nbI; }
      sIter = i;
    }
    __syncthreads();
    // compact the live list once a quarter of it has died since the last rebuild
    const int it = sIter;
    if ((it - iterAtRebuild) * 4 > liveAtRebuild) {
      liveLen = rebuild_live(S, maxbins, live, posOf, sWarp);
      liveAtRebuild = liveLen; iterAtRebuild = it;
    }
  }

  // ---- palette fill (PQ:258-264, PL:315-324): the k-th live bin in ascending order
  liveLen = rebuild_live(S, maxbins, live, posOf, sWarp);
  const int plen = extbins > 0 ? I.nmax : maxbins;
  for (int k = t; k < plen; k += NQ_MERGE_THREADS) {
    const int b = live[k];
    uint32_t colr;
    if (rgb) colr = c_argb(j2i(S.bAc[b]), j2i(S.bC1[b]), j2i(S.bC2[b]), j2i(S.bC3[b]));
    else {
      if (!lab2rgb(j2i((double)S.fAc[b]), S.fC1[b], S.fC2[b], S.fC3[b], &colr)) { colr = 0; I.error = 3; }
    }
    I.palette[k] = colr;
  }
  if (t == 0) {
    I.paletteLen = plen;
    I.statRescans = rescans; I.statPairs += pairs; I.statHeapPops = pops;
  }
}

}  // namespace nq
