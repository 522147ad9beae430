// periodic_index.h -- stage-periodic compression of the index structures of an OCP-shaped QP.
//
// The CCS pattern of a multiple-shooting OCP repeats from stage to stage: the columns of stage
// k + 1 hold the rows of stage k shifted by a constant per entry (identity rows move by the frame
// size, dynamics rows by the state size, ...), and the value positions move by the non-zeros of one
// stage.  Instead of one 16-bit index per non-zero (22 KB of shared memory per CTA for the H = 20
// quadrotor, more than the matrix values themselves) the compact throughput kernel
// (admm_compact_kernel.cuh) keeps ONE template per run of identical stages:
//
//   outer index space (columns of a CSC structure, rows of a CSR one) = a few REGIONS
//     explicit region   every outer index has its own template slot (reps == 1)
//     periodic region   `reps` repetitions of `period` outer indices; repetition q of template slot t
//                       covers outer index i0 + q * period + (t - t0), its entries sit at template
//                       positions tptr[t] .. tptr[t + 1] and mean
//                           value    = (tent[kt] & 0xffff) + q * (tent[kt] >> 16)
//                           position = kt + koff + q * dk          (position in the uncompressed arrays)
//
// The builder finds the regions by itself (no knowledge of the problem beyond a list of candidate
// periods), verifies the result by expanding it again, and reports failure instead of guessing:
// patterns without stage structure simply end up as one explicit region, which is never smaller than
// the plain arrays and makes the caller keep the uncompressed kernel.
#pragma once

#include <stdint.h>

#ifndef __CUDACC__
#define OCP_B200_HD
#else
#define OCP_B200_HD __host__ __device__ __forceinline__
#endif

namespace ocpb200 {

constexpr int kMaxRegions = 12;

struct PRegion {
  int i0, i1;        // outer index range [i0, i1)
  int period;        // outer indices per repetition; i1 - i0 == period * reps
  int t0;            // template slot of outer index i0
  int koff;          // position = template position + koff + q * dk
  int dk;            // positions per repetition
  int k0, k1;        // position range [k0, k1) of the region in the uncompressed arrays
  int kt0;           // template position of k0
  float inv_period;  // 1 / period
  float inv_dk;      // 1 / dk  (0 when dk == 0)
  int uni;           // (d + 1) | (d2 + 1) << 16 when every entry of the region moves by the same delta d (first
                     // payload) / d2 (second payload) per repetition, 0 in a half that is not uniform
};

// one compressed structure; all pointers address 32-bit words of one arena
struct PIndex {
  int nreg;
  const PRegion* reg;
  const uint32_t* tptr;   // template pointers (nt + 1)
  const uint32_t* tent;   // template entries: value | delta << 16
  const uint32_t* tent2;  // second payload (CSR of A: position in the CSC value array), or null
};

struct PSpan {
  int kt0, kt1;   // template positions of the entries
  int q;          // repetition
  int kshift;     // position in the uncompressed arrays = template position + kshift
  int qd, qd2;    // device spans: q * (uniform delta) of the two payloads, -1 when the region's deltas differ
};

// low half of a template entry (the value of repetition 0) read as a 16-bit load
OCP_B200_HD int plow(const uint32_t* tent, int kt) { return reinterpret_cast<const uint16_t*>(tent)[2 * kt]; }

// entries of outer index i
OCP_B200_HD PSpan pspan(const PIndex& X, int i) {
  int r = 0;
  while (r + 1 < X.nreg && i >= X.reg[r].i1) ++r;
  const PRegion R = X.reg[r];
  const int o = i - R.i0;
  const int q = R.period < R.i1 - R.i0 ? static_cast<int>((static_cast<float>(o) + 0.5f) * R.inv_period) : 0;
  const int t = R.t0 + o - q * R.period;
  PSpan s;
  s.kt0 = static_cast<int>(X.tptr[t]);
  s.kt1 = static_cast<int>(X.tptr[t + 1]);
  s.q = q;
  s.kshift = R.koff + q * R.dk;
  s.qd = (R.uni & 0xffff) ? q * ((R.uni & 0xffff) - 1) : -1;
  s.qd2 = (R.uni >> 16) ? q * ((R.uni >> 16) - 1) : -1;
  return s;
}
OCP_B200_HD int pvalue(uint32_t ent, int q) { return static_cast<int>(ent & 0xffffu) + q * static_cast<int>(ent >> 16); }

// value of the entry at uncompressed position k (random access; K assembly only)
OCP_B200_HD int pvalue_at(const PIndex& X, int k) {
  int r = 0;
  while (r + 1 < X.nreg && k >= X.reg[r].k1) ++r;
  const PRegion R = X.reg[r];
  const int o = k - R.k0;
  const int q = (R.dk > 0 && R.period < R.i1 - R.i0) ? static_cast<int>((static_cast<float>(o) + 0.5f) * R.inv_dk) : 0;
  return pvalue(X.tent[R.kt0 + o - q * R.dk], q);
}

// Device view: the region boundaries travel as kernel parameters (constant bank: the region search is a
// few compares against immediates, no dependent loads), everything else sits in the 32-bit arena that
// every CTA copies into shared memory.
constexpr int kMaxDevRegions = 6;
struct PIndexDev {
  int nreg;
  int ibound[kMaxDevRegions];   // i1 of every region (outer index space)
  int kbound[kMaxDevRegions];   // k1 of every region (position space)
  int reg_off, tptr_off, tent_off, tent2_off;   // word offsets in the arena (tent2_off < 0: none)
};

// compressed index structures of one pattern (kernel parameter of the compact throughput kernel)
struct CompactIdx {
  const uint32_t* arena;   // global copy of the 32-bit arena
  int arena_words;
  PIndexDev acol;          // CSC of A: entries = row indices
  PIndexDev arow;          // CSR of A: entries = column indices, second payload = position in the CSC values
  PIndexDev pcol;          // symmetrised P, CSC: entries = row indices
};

#ifdef __CUDACC__
__device__ __forceinline__ const PRegion* pregion(const PIndexDev& C, const uint32_t* ar, const int (&bound)[kMaxDevRegions], int v) {
  int r = 0;
#pragma unroll
  for (int t = 0; t + 1 < kMaxDevRegions; ++t) r += (t + 1 < C.nreg && v >= bound[t]) ? 1 : 0;
  return reinterpret_cast<const PRegion*>(ar + C.reg_off) + r;
}
__device__ __forceinline__ PSpan pspan_dev(const PIndexDev& C, const uint32_t* ar, int i) {
  const PRegion* R = pregion(C, ar, C.ibound, i);
  const int i0 = R->i0, i1 = R->i1, period = R->period, t0 = R->t0, koff = R->koff, dk = R->dk, uni = R->uni;
  const float inv = R->inv_period;
  const int o = i - i0;
  const int q = period < i1 - i0 ? static_cast<int>((static_cast<float>(o) + 0.5f) * inv) : 0;
  const uint32_t* tp = ar + C.tptr_off + t0 + o - q * period;
  PSpan s;
  s.kt0 = static_cast<int>(tp[0]);
  s.kt1 = static_cast<int>(tp[1]);
  s.q = q;
  s.kshift = koff + q * dk;
  s.qd = (uni & 0xffff) ? q * ((uni & 0xffff) - 1) : -1;
  s.qd2 = (uni >> 16) ? q * ((uni >> 16) - 1) : -1;
  return s;
}
__device__ __forceinline__ int pvalue_at_dev(const PIndexDev& C, const uint32_t* ar, int k) {
  const PRegion* R = pregion(C, ar, C.kbound, k);
  const int o = k - R->k0, dk = R->dk;
  const int q = (dk > 0 && R->period < R->i1 - R->i0) ? static_cast<int>((static_cast<float>(o) + 0.5f) * R->inv_dk) : 0;
  return pvalue(ar[C.tent_off + R->kt0 + o - q * dk], q);
}
#endif

}  // namespace ocpb200

#include <algorithm>
#include <vector>

namespace ocpb200 {

// host-side result: regions + template arrays, laid out later into one arena
struct PIndexHost {
  std::vector<PRegion> reg;
  std::vector<uint32_t> tptr, tent, tent2;
  size_t words() const { return reg.size() * (sizeof(PRegion) / 4) + tptr.size() + tent.size() + tent2.size(); }
};

// Compresses a CSC/CSR-like structure: ptr (nout + 1), val (first payload, e.g. row indices), val2
// (optional second payload of the same length, may be empty).  `periods`: candidate repetition lengths
// in outer indices.  Returns false when the structure cannot be represented (values or deltas beyond
// 16 bits, more than kMaxRegions regions, or the self-check fails).
inline bool build_periodic_index(const std::vector<int>& ptr, const std::vector<int>& val, const std::vector<int>& val2,
                                 std::vector<int> periods, PIndexHost& out) {
  const int nout = static_cast<int>(ptr.size()) - 1;
  const bool two = !val2.empty();
  out = PIndexHost();
  std::sort(periods.begin(), periods.end());
  periods.erase(std::unique(periods.begin(), periods.end()), periods.end());
  // how many repetitions of `p` outer indices starting at i have the same shape with constant deltas?
  auto reps_at = [&](int i, int p) {
    if (p <= 0 || i + 2 * p > nout) return 1;
    const int len = ptr[i + p] - ptr[i];
    int reps = 1;
    while (i + (reps + 1) * p <= nout) {
      const int a = i + (reps - 1) * p, b = i + reps * p;   // repetition reps-1 vs reps
      bool ok = ptr[b + p] - ptr[b] == len;
      for (int c = 0; ok && c < p; ++c) ok = ptr[a + c + 1] - ptr[a + c] == ptr[b + c + 1] - ptr[b + c];
      for (int e = 0; ok && e < len; ++e) {
        const int d = val[ptr[b] + e] - val[ptr[a] + e], d0 = val[ptr[i + p] + e] - val[ptr[i] + e];
        ok = d == d0 && d >= 0 && d < 65536;
        if (ok && two) {
          const int e2 = val2[ptr[b] + e] - val2[ptr[a] + e], e20 = val2[ptr[i + p] + e] - val2[ptr[i] + e];
          ok = e2 == e20 && e2 >= 0 && e2 < 65536;
        }
      }
      if (!ok) break;
      ++reps;
    }
    return reps;
  };
  struct Cut { int i0, period, reps; };
  // segmentation by dynamic programming over the outer index: f[i][e] = fewest words for [i, nout) when the
  // index before i is (e = 1) / is not (e = 0) part of an explicit region.  Every region costs its descriptor
  // plus a penalty that keeps the number of regions small.
  const int wpe = two ? 2 : 1, kRegionCost = static_cast<int>(sizeof(PRegion) / 4) + 52;
  std::vector<int> best_p(nout + 1, 0), best_r(nout + 1, 1), per_cost(nout + 1, 0);
  std::vector<long> f0(nout + 1, 0), f1(nout + 1, 0);
  std::vector<char> c0(nout + 1, 0), c1(nout + 1, 0);   // 1: start a periodic region at i
  for (int i = nout - 1; i >= 0; --i) {
    // best periodic region starting at i (all its repetitions)
    long bestv = -1;
    for (int p : periods) {
      const int reps = reps_at(i, p);
      if (reps < 2) continue;
      const long v = kRegionCost + p + long(ptr[i + p] - ptr[i]) * wpe + f0[i + p * reps];
      if (bestv < 0 || v < bestv) { bestv = v; best_p[i] = p; best_r[i] = reps; }
    }
    const long ce = 1 + long(ptr[i + 1] - ptr[i]) * wpe;
    f1[i] = ce + f1[i + 1]; c1[i] = 0;
    if (bestv >= 0 && bestv < f1[i]) { f1[i] = bestv; c1[i] = 1; }
    f0[i] = kRegionCost + ce + f1[i + 1]; c0[i] = 0;
    if (bestv >= 0 && bestv < f0[i]) { f0[i] = bestv; c0[i] = 1; }
  }
  std::vector<Cut> cuts;
  {
    int i = 0, explicit_from = -1;
    while (i < nout) {
      const bool periodic = explicit_from >= 0 ? c1[i] : c0[i];
      if (periodic) {
        if (explicit_from >= 0) { cuts.push_back({explicit_from, i - explicit_from, 1}); explicit_from = -1; }
        cuts.push_back({i, best_p[i], best_r[i]});
        i += best_p[i] * best_r[i];
      } else {
        if (explicit_from < 0) explicit_from = i;
        ++i;
      }
    }
    if (explicit_from >= 0) cuts.push_back({explicit_from, nout - explicit_from, 1});
  }
  if (cuts.empty()) cuts.push_back({0, 0, 1});
  if (static_cast<int>(cuts.size()) > kMaxRegions) return false;
  for (const Cut& c : cuts) {
    PRegion R{};
    R.i0 = c.i0; R.i1 = c.i0 + c.period * c.reps; R.period = c.period;
    R.t0 = static_cast<int>(out.tptr.size());
    R.k0 = ptr[R.i0]; R.k1 = ptr[R.i1];
    R.kt0 = static_cast<int>(out.tent.size());
    R.dk = c.reps > 1 ? ptr[c.i0 + c.period] - ptr[c.i0] : 0;
    R.koff = R.k0 - R.kt0;
    R.inv_period = c.period > 0 ? 1.0f / static_cast<float>(c.period) : 0.0f;
    R.inv_dk = R.dk > 0 ? 1.0f / static_cast<float>(R.dk) : 0.0f;
    int ud = -2, ud2 = -2;   // common delta of the region's entries: -2 none seen yet, -1 not uniform
    for (int o = 0; o < c.period; ++o) {
      out.tptr.push_back(static_cast<uint32_t>(out.tent.size()));
      for (int k = ptr[c.i0 + o]; k < ptr[c.i0 + o + 1]; ++k) {
        const int d = c.reps > 1 ? val[k + R.dk] - val[k] : 0;
        if (val[k] < 0 || val[k] >= 65536) return false;
        out.tent.push_back(static_cast<uint32_t>(val[k]) | (static_cast<uint32_t>(d) << 16));
        ud = ud == -2 ? d : (ud == d ? ud : -1);
        if (two) {
          const int d2 = c.reps > 1 ? val2[k + R.dk] - val2[k] : 0;
          if (val2[k] < 0 || val2[k] >= 65536) return false;
          out.tent2.push_back(static_cast<uint32_t>(val2[k]) | (static_cast<uint32_t>(d2) << 16));
          ud2 = ud2 == -2 ? d2 : (ud2 == d2 ? ud2 : -1);
        }
      }
    }
    if (ud == -2) ud = 0;
    if (ud2 == -2) ud2 = 0;
    R.uni = ((ud >= 0 && ud < 65535) ? ud + 1 : 0) | (((ud2 >= 0 && ud2 < 65535) ? ud2 + 1 : 0) << 16);
    out.reg.push_back(R);
  }
  out.tptr.push_back(static_cast<uint32_t>(out.tent.size()));   // one shared end pointer: slots are consecutive
  // self-check: expand and compare, both through the outer index and through the position
  PIndex X{static_cast<int>(out.reg.size()), out.reg.data(), out.tptr.data(), out.tent.data(), two ? out.tent2.data() : nullptr};
  for (int o = 0; o < nout; ++o) {
    const PSpan s = pspan(X, o);
    if (s.kt1 - s.kt0 != ptr[o + 1] - ptr[o] || s.kt0 + s.kshift != ptr[o]) return false;
    for (int kt = s.kt0; kt < s.kt1; ++kt) {
      if (s.qd >= 0 && plow(X.tent, kt) + s.qd != val[kt + s.kshift]) return false;
      if (two && s.qd2 >= 0 && plow(X.tent2, kt) + s.qd2 != val2[kt + s.kshift]) return false;
      if (pvalue(X.tent[kt], s.q) != val[kt + s.kshift]) return false;
      if (two && pvalue(X.tent2[kt], s.q) != val2[kt + s.kshift]) return false;
      if (pvalue_at(X, kt + s.kshift) != val[kt + s.kshift]) return false;
    }
  }
  return true;
}

}  // namespace ocpb200
