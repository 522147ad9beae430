// tri_fast.cuh -- exact-size (compile-time BS) versions of the bordered block-tridiagonal LDL'
// factorisation and solve of admm_direct_kernel.cuh.
//
// The sequential parts of the algorithm (the two sweeps of a solve, the Gauss-Jordan inverse
// of a diagonal block) run in ONE warp while the rest of the CTA waits, so what limits them is
// the length of that warp's instruction stream and its dependent-latency chain -- not memory
// bandwidth.  These versions therefore: unroll everything over BS with no predication, read
// block rows with 16-byte shared-memory loads (row pitch ld is even), split every block row
// over LPR = 32 / BS lanes (BS = 16: two lanes per row, one shuffle to combine), keep the next
// stage's block row in registers while the current one is multiplied, and hold one matrix
// column per lane in registers during the inverse.  No integer division in per-stage loops.
#pragma once

namespace ocpb200 {
namespace direct {

template <int BS> struct TriCfg {
  static constexpr int LPR = BS <= 8 ? 4 : (BS <= 16 ? 2 : 1);   // lanes per block row in the sweeps
  static constexpr int CPL = BS / LPR;                           // columns per lane
  static_assert(BS % LPR == 0 && CPL % 2 == 0, "block size must split into even lane chunks");
};

// sum_t a[t] * b[t],  both contiguous and 16-byte aligned, t < BS
template <int BS>
__device__ __forceinline__ double dot_cc(const double* __restrict__ a, const double* __restrict__ b) {
  const double2* a2 = reinterpret_cast<const double2*>(a);
  const double2* b2 = reinterpret_cast<const double2*>(b);
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
  for (int t = 0; t < BS / 2; t += 2) {
    const double2 x0 = a2[t], y0 = b2[t];
    s0 = fma(x0.x, y0.x, s0); s1 = fma(x0.y, y0.y, s1);
    if (t + 1 < BS / 2) {
      const double2 x1 = a2[t + 1], y1 = b2[t + 1];
      s2 = fma(x1.x, y1.x, s2); s3 = fma(x1.y, y1.y, s3);
    }
  }
  return (s0 + s1) + (s2 + s3);
}

// sum_t a[t] * b[t * sb],  a contiguous and 16-byte aligned, b strided
template <int BS>
__device__ __forceinline__ double dot_cs(const double* __restrict__ a, const double* __restrict__ b, int sb) {
  const double2* a2 = reinterpret_cast<const double2*>(a);
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
  for (int t = 0; t < BS / 2; t += 2) {
    const double2 x0 = a2[t];
    s0 = fma(x0.x, b[(2 * t) * sb], s0); s1 = fma(x0.y, b[(2 * t + 1) * sb], s1);
    if (t + 1 < BS / 2) {
      const double2 x1 = a2[t + 1];
      s2 = fma(x1.x, b[(2 * t + 2) * sb], s2); s3 = fma(x1.y, b[(2 * t + 3) * sb], s3);
    }
  }
  return (s0 + s1) + (s2 + s3);
}

// in-place inverse of an SPD BS x BS block (Gauss-Jordan, no pivoting) by ONE warp.
// Lane c < BS holds column c in registers.  Per pivot k the multipliers f[r] = -M[r][k] / M[k][k]
// are needed by every lane; M[r][k] is (up to the sign flip of already pivoted rows) element k of
// lane r's own column, because the working matrix stays symmetric among unpivoted indices:
// lane k broadcasts 1 / M[k][k] (one shuffle), every lane forms its own f[r] with one multiply and
// the 16 values are all-gathered through `piv` (16-byte aligned scratch of >= BS doubles).  The
// serial part of a pivot is reciprocal -> shuffle -> multiply -> store/load round trip -> FMAs.
template <int BS, int LD = BS + 2>
__device__ __forceinline__ void warp_invert_exact(double* M, int /*ld*/, int lane, double* piv) {
  constexpr int ld = LD;
  static_assert(BS % 2 == 0 && BS <= 32, "one column per lane, even size");
  const bool act = lane < BS;
  const int cc = act ? lane : 0;
  double col[BS];
#pragma unroll
  for (int r = 0; r < BS; ++r) col[r] = M[r * ld + cc];
  const double2* piv2 = reinterpret_cast<const double2*>(piv);
#pragma unroll
  for (int k = 0; k < BS; ++k) {
    const double ck = col[k];                                   // M[k][c] == M[c][k]
    const double ipiv = 1.0 / __shfl_sync(0xffffffffu, ck, k);  // every lane: 1 / M[k][k]
    const bool mine = lane == k;
    // f[lane] = -M[lane][k] / M[k][k].  Rows already pivoted hold the sign-flipped coupling
    // (M[r][k] == -M[k][r] for r < k in in-place Gauss-Jordan), rows still to come are symmetric.
    if (act) piv[lane] = mine ? ipiv : (lane < k ? ck * ipiv : -ck * ipiv);
    __syncwarp();
#pragma unroll
    for (int r = 0; r < BS / 2; ++r) {
      const double2 f = piv2[r];
      const int r0 = 2 * r, r1 = 2 * r + 1;
      col[r0] = mine ? f.x : (r0 == k ? ck * f.x : fma(f.x, ck, col[r0]));
      col[r1] = mine ? f.y : (r1 == k ? ck * f.y : fma(f.y, ck, col[r1]));
    }
    __syncwarp();
  }
  if (act) {
#pragma unroll
    for (int r = 0; r < BS; ++r) M[r * ld + lane] = col[r];
  }
}

// D_p^-1 (np x np, pitch np + 1): exact-size code for the border sizes of the benchmark problems
__device__ __forceinline__ void invert_border(double* Dp, int np, int lane, double* piv) {
  if (np == 12) warp_invert_exact<12, 13>(Dp, 13, lane, piv);
  else if (np == 4) warp_invert_exact<4, 5>(Dp, 5, lane, piv);
  else if (np == 24) warp_invert_exact<24, 25>(Dp, 25, lane, piv);
  else warp_invert(Dp, np, np + 1, lane);
}

template <int BS>
__device__ inline void tri_factor_exact(const PatternDev& P, const Work& W) {
  const int tid = threadIdx.x, T = blockDim.x, lane = tid & 31, warp = tid >> 5;
  const int np = P.tri_np, nb = P.tri_nb, ld = P.tri_ld, N = nb * BS;
  constexpr int bb = BS * BS;
  const int pb = np * BS;
  const int r1 = tid / BS, c1 = tid % BS;
  const int rstep = T / BS, cstep = T % BS;   // advancing an element index by T
  for (int k = 0; k < nb; ++k) {
    double* Dk = W.Dinv + size_t(k) * BS * ld;
    if (k > 0) {
      double* Tk = W.Lsub + size_t(k) * BS * ld;               // K_{k,k-1}
      const double* Dm = W.Dinv + size_t(k - 1) * BS * ld;     // D_{k-1}^-1 (symmetric)
      // S = K_{k,k-1} D_{k-1}^-1  (= L_k; D^-1 symmetric: its row c is its column c);  Sp = V_{k-1}
      for (int e = tid, r = r1, c = c1; e < bb; e += T) {
        W.S[r * ld + c] = dot_cc<BS>(Tk + r * ld, Dm + c * ld);
        r += rstep; c += cstep; if (c >= BS) { c -= BS; ++r; }
      }
      for (int e = tid, r = r1, c = c1; e < pb; e += T) {
        W.Sp[e] = W.Lp[size_t(r) * N + (k - 1) * BS + c];
        r += rstep; c += cstep; if (c >= BS) { c -= BS; ++r; }
      }
      __syncthreads();
      // D_k = K_kk - S K_{k,k-1}';  V_k = K_pk - V_{k-1} S';  L_{p,k-1} = V_{k-1} D_{k-1}^-1
      for (int e = tid, r = r1, c = c1; e < bb; e += T) {
        Dk[r * ld + c] -= dot_cc<BS>(W.S + r * ld, Tk + c * ld);
        r += rstep; c += cstep; if (c >= BS) { c -= BS; ++r; }
      }
      for (int e = tid, r = r1, c = c1; e < pb; e += T) {
        const double s = dot_cc<BS>(W.Sp + r * BS, W.S + c * ld);
        const double f = dot_cc<BS>(W.Sp + r * BS, Dm + c * ld);
        W.Lp[size_t(r) * N + k * BS + c] -= s;
        W.Lp[size_t(r) * N + (k - 1) * BS + c] = f;
        r += rstep; c += cstep; if (c >= BS) { c -= BS; ++r; }
      }
      __syncthreads();
    }
    if (warp == 0) {
      warp_invert_exact<BS>(Dk, ld, lane, W.piv);
    } else if (k > 0) {
      // the other warps, meanwhile: D_p -= V_{k-1} L_{p,k-1}';  L_k <- S
      double* Tk = W.Lsub + size_t(k) * BS * ld;
      const int t2 = tid - 32, T2 = T - 32;
      for (int e = t2; e < np * np; e += T2) {
        const int r = e / np, c = e - r * np;
        W.Dp[r * (np + 1) + c] -= dot_cs<BS>(W.Sp + r * BS, W.Lp + size_t(c) * N + (k - 1) * BS, 1);
      }
      for (int e = t2; e < bb; e += T2) {
        const int r = e / BS, c = e % BS;
        Tk[r * ld + c] = W.S[r * ld + c];
      }
    }
    __syncthreads();
  }
  // last border block, then D_p^-1
  if (np > 0) {
    const double* Dm = W.Dinv + size_t(nb - 1) * BS * ld;
    for (int e = tid, r = r1, c = c1; e < pb; e += T) {
      W.Sp[e] = W.Lp[size_t(r) * N + (nb - 1) * BS + c];
      r += rstep; c += cstep; if (c >= BS) { c -= BS; ++r; }
    }
    __syncthreads();
    for (int e = tid, r = r1, c = c1; e < pb; e += T) {
      W.Lp[size_t(r) * N + (nb - 1) * BS + c] = dot_cc<BS>(W.Sp + r * BS, Dm + c * ld);
      r += rstep; c += cstep; if (c >= BS) { c -= BS; ++r; }
    }
    __syncthreads();
    for (int e = tid; e < np * np; e += T) {
      const int r = e / np, c = e - r * np;
      W.Dp[r * (np + 1) + c] -= dot_cs<BS>(W.Sp + r * BS, W.Lp + size_t(c) * N + (nb - 1) * BS, 1);
    }
    __syncthreads();
    if (warp == 0) warp_invert(W.Dp, np, np + 1, lane);
    __syncthreads();
  }
}

// one stage of a sweep: dst[r] -= sum_c Lpart[c] * src[c] over this lane's CPL columns, combined
// over the LPR lanes of the row
template <int BS>
__device__ __forceinline__ void sweep_stage(const double (&Lpart)[TriCfg<BS>::CPL], const double* __restrict__ src,
                                            double* __restrict__ dst, bool writer) {
  constexpr int CPL = TriCfg<BS>::CPL, LPR = TriCfg<BS>::LPR;
  const double old = writer ? *dst : 0.0;   // issued before the chain, not after the shuffle
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
  for (int c = 0; c < CPL; c += 4) {
    s0 = fma(Lpart[c], src[c], s0);
    s1 = fma(Lpart[c + 1], src[c + 1], s1);
    if (c + 2 < CPL) { s2 = fma(Lpart[c + 2], src[c + 2], s2); s3 = fma(Lpart[c + 3], src[c + 3], s3); }
  }
  double s = (s0 + s1) + (s2 + s3);
  if (LPR >= 2) s += __shfl_xor_sync(0xffffffffu, s, 1);
  if (LPR >= 4) s += __shfl_xor_sync(0xffffffffu, s, 2);
  if (writer) *dst = old - s;
  __syncwarp();
}

// K x = b in place: b is [p | block 0 | ... | block nb-1]
template <int BS>
__device__ inline void tri_solve_exact(const PatternDev& P, const Work& W) {
  constexpr int CPL = TriCfg<BS>::CPL, LPR = TriCfg<BS>::LPR;
  const int tid = threadIdx.x, T = blockDim.x, lane = tid & 31, warp = tid >> 5, nw = T >> 5;
  const int np = P.tri_np, nb = P.tri_nb, ld = P.tri_ld, N = nb * BS;
  double* bx = W.b + np;
  const int row = lane / LPR, half = lane % LPR;
  const bool act = row < BS;
  const int rr = act ? row : 0;
  const bool writer = act && half == 0;
  // forward sweep: y_k = b_k - L_k y_{k-1}   (one warp; next stage's row prefetched in registers)
  if (warp == 0 && nb > 1) {
    auto load_row = [&](double (&dst)[CPL], int k) {
      const double2* p = reinterpret_cast<const double2*>(W.Lsub + size_t(k) * BS * ld + rr * ld + half * CPL);
#pragma unroll
      for (int i = 0; i < CPL / 2; ++i) { const double2 v = p[i]; dst[2 * i] = v.x; dst[2 * i + 1] = v.y; }
    };
    double La[CPL], Lb[CPL];
    load_row(La, 1);
    for (int k = 1; k < nb; k += 2) {
      if (k + 1 < nb) load_row(Lb, k + 1);
      sweep_stage<BS>(La, bx + (k - 1) * BS + half * CPL, bx + k * BS + rr, writer);
      if (k + 1 < nb) {
        if (k + 2 < nb) load_row(La, k + 2);
        sweep_stage<BS>(Lb, bx + k * BS + half * CPL, bx + (k + 1) * BS + rr, writer);
      }
    }
  }
  __syncthreads();
  // border: y_p = b_p - sum_k L_pk y_k  (a warp per border row), x_p = D_p^-1 y_p
  if (np > 0) {
    for (int r = warp; r < np; r += nw) {
      const double* rowp = W.Lp + size_t(r) * N;
      double s0 = 0.0, s1 = 0.0;
      int j = lane;
      for (; j + 32 < N; j += 64) { s0 = fma(rowp[j], bx[j], s0); s1 = fma(rowp[j + 32], bx[j + 32], s1); }
      if (j < N) s0 = fma(rowp[j], bx[j], s0);
      double s = s0 + s1;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (lane == 0) W.xp[r] = W.b[r] - s;
    }
    __syncthreads();
    if (tid < np) {
      double s = 0.0;
      for (int c = 0; c < np; ++c) s = fma(W.Dp[tid * (np + 1) + c], W.xp[c], s);
      W.b[tid] = s;
    }
    __syncthreads();
  }
  // diagonal: c_k = D_k^-1 y_k - L_pk' x_p   (in place: value computed, barrier, stored)
  {
    const int Tb = (T / BS) * BS;   // whole blocks per pass: a pass never reads what it overwrites
    const int k1 = tid / BS, r1 = tid % BS, kstep = T / BS;
    for (int base = 0, k = k1; base < N; base += Tb, k += kstep) {
      const int j = tid < Tb ? base + tid : N;
      double v = 0.0;
      if (j < N) {
        v = dot_cs<BS>(W.Dinv + size_t(k) * BS * ld + r1 * ld, bx + k * BS, 1);
        double v1 = 0.0;
        int p = 0;
        for (; p + 1 < np; p += 2) {
          v = fma(-W.Lp[size_t(p) * N + j], W.b[p], v);
          v1 = fma(-W.Lp[size_t(p + 1) * N + j], W.b[p + 1], v1);
        }
        if (p < np) v = fma(-W.Lp[size_t(p) * N + j], W.b[p], v);
        v += v1;
      }
      __syncthreads();
      if (j < N) bx[j] = v;
    }
  }
  __syncthreads();
  // backward sweep: x_k = c_k - L_{k+1}' x_{k+1}   (lane (r, half) reads part of column r of L_{k+1})
  if (warp == 0 && nb > 1) {
    auto load_col = [&](double (&dst)[CPL], int k) {
      const double* p = W.Lsub + size_t(k) * BS * ld + size_t(half * CPL) * ld + rr;
#pragma unroll
      for (int i = 0; i < CPL; ++i) dst[i] = p[i * ld];
    };
    double La[CPL], Lb[CPL];
    load_col(La, nb - 1);
    for (int k = nb - 2; k >= 0; k -= 2) {
      if (k > 0) load_col(Lb, k);
      sweep_stage<BS>(La, bx + (k + 1) * BS + half * CPL, bx + k * BS + rr, writer);
      if (k > 0) {
        if (k > 1) load_col(La, k - 1);
        sweep_stage<BS>(Lb, bx + k * BS + half * CPL, bx + (k - 1) * BS + rr, writer);
      }
    }
  }
  __syncthreads();
}

}  // namespace direct
}  // namespace ocpb200
