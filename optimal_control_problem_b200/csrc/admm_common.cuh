// admm_common.cuh -- pieces shared by the two batched ADMM QP kernels for sm_100a
// (admm_direct_kernel.cuh: bordered block-tridiagonal LDL' of the reduced KKT matrix;
//  admm_pcg_kernel.cuh: preconditioned CG on the same matrix for patterns without that structure).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "ocp_b200.h"

namespace ocpb200 {

constexpr double kInfty = 1e30, kMinScaling = 1e-4, kMaxScaling = 1e4;
constexpr double kRhoMin = 1e-6, kRhoMax = 1e6, kRhoTol = 1e-4, kRhoEqOverIneq = 1e3;
constexpr double kDivisionTol = 1e-30;
constexpr int kRedWidth = 16;          // max values per block reduction
constexpr int kMaxWarps = 32;

typedef uint16_t idx_t;

// -DOCP_B200_CANARY: every per-instance array (shared memory and slab) is followed by two guard doubles that the
// kernels fill when they carve their state and verify when they leave; an overwritten guard is reported with
// printf ("OCP_B200 CANARY ...").  compute-sanitizer is closed on the GPU pool this was developed on: this is the
// bounds check of our own (tools/sanitize_cases.py runs every placement with it, profiles/r2_canary.txt).
#ifdef OCP_B200_CANARY
constexpr int kCanaryDoubles = 2;
#else
constexpr int kCanaryDoubles = 0;
#endif
__device__ __forceinline__ double canary_value(int id) { return __longlong_as_double(0x7ff4dead00000000LL + id); }

// K-assembly program (direct kernel): one entry per structurally non-zero element of the bordered
// block-tridiagonal K = P + sigma I + A' diag(rho) A that has to be computed, with the products
// sum_r rho_r A_ri A_rj encoded as RUNS of consecutive positions in columns i and j of A whose rows
// coincide one-to-one AND are consecutive rows row0, row0 + 1, ... (host-side merge of the two sorted
// columns, done once per pattern), so that the constraint type of every product needs no index lookup.
struct KRun { uint16_t ka, kc, len, row0; };
struct KEntry {
  uint32_t dest0, dest1;   // (array << 30) | offset; array 0 = Dinv, 1 = Lsub, 2 = Lp, 3 = Dp; dest1 = mirror or 0xffffffff
  int32_t ppos;            // position of P_ij in the symmetrised P values, or -1
  uint32_t run_begin;      // first run in the run array
  uint16_t nruns, diag;    // diag != 0: i == j (sigma is added)
  uint32_t pad;
};

// Index structures of one sparsity pattern (device pointers, shared by every instance)
struct PatternDev {
  int n, m, nnz_a, nnz_p, nnz_h;
  int nblk, minv_doubles, max_bs;
  int n_long, n_short;
  // bordered block-tridiagonal structure of K = P + sigma I + A' diag(rho) A (direct kernel):
  // columns [0, tri_np) form the border (the reference parameters p), then tri_nb diagonal
  // blocks of tri_bs columns each; block rows are stored with an even pitch tri_ld >= tri_bs + 1
  int tri_ok, tri_np, tri_bs, tri_nb, tri_ld;
  int stage_slots;         // block staging buffers per CTA when the factor is slab-resident (4, or 8 = a ring of 4 per chain)
  int kprog_entries;       // K-assembly program (null / 0: assemble by merging columns on the device)
  const KEntry* kprog;
  const KRun* kruns;
  int idx_entries;         // length of the idx_t arena (every array padded to 8 entries)
  const idx_t* idx_base;   // start of the arena
  const idx_t* a_colptr;   // n+1
  const idx_t* a_rowidx;   // nnz_a
  const idx_t* a_rowptr;   // m+1
  const idx_t* a_colidx;   // nnz_a, CSR order
  const idx_t* a_perm;     // nnz_a, CSR position -> CSC position
  const idx_t* p_colptr;   // n+1   (symmetrised full pattern of the upper triangle of H)
  const idx_t* p_rowidx;   // nnz_p
  const int* p_src;        // nnz_p, index into the caller's H values (upper-triangle twin)
  const idx_t* blk_ptr;    // nblk+1
  const idx_t* blk_of_col; // n
  const int* minv_off;     // nblk
  const idx_t* rows_long;  // rows handled by 4 lanes each
  const idx_t* rows_short; // rows handled by one thread each
};

struct SolveArgs {
  int B;
  // QP data, one row per instance
  const double* h_vals; int ld_h;
  const double* q; int ld_n;
  const double* a_vals; int ld_a;
  const double* l; const double* u; int ld_m;
  // outputs
  double* sol_x;      // B*n   unscaled primal solution (may be null)
  double* sol_y;      // B*m   unscaled dual solution (may be null)
  double* info;       // B*OCP_B200_NINFO (may be null)
  // SQP update: x_iter[b*N + i] += alpha * sol[np + i]; stats accumulated (may be null)
  double* x_iter; int np; int N; double sqp_alpha;
  double* stats; int first_step;
  // trace of instance 0 (may be null)
  double* trace; int max_trace; int* n_trace;
  // optional phase cycle counters of CTA 0 (OCP_B200_NPHASE long longs, may be null)
  long long* phase;
  // scheduling + streaming workspace
  int* counter;
  double* slab; size_t slab_doubles;
};

// ---------------------------------------------------------------------------------------
// block reductions: warp shuffles, then one shared-memory exchange; every thread returns
// with the same result.  Two alternating exchange buffers make one barrier per call enough.
// ---------------------------------------------------------------------------------------
struct Reducer {
  double* buf;   // 2 * half doubles of shared memory, half >= (warps per CTA) * kRedWidth
  int parity;
  int half;
};

template <int NV, bool kMax>
__device__ __forceinline__ void block_reduce(double (&v)[NV], Reducer& R) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    double a = v[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double b = __shfl_xor_sync(0xffffffffu, a, o);
      a = kMax ? fmax(a, b) : a + b;
    }
    v[k] = a;
  }
  double* buf = R.buf + R.parity * R.half;
  R.parity ^= 1;
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < NV; ++k) buf[warp * NV + k] = v[k];
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    double a = buf[k];
    for (int w = 1; w < nw; ++w) a = kMax ? fmax(a, buf[w * NV + k]) : a + buf[w * NV + k];
    v[k] = a;
  }
}

__device__ __forceinline__ double limit_scaling(double v) {
  v = v < kMinScaling ? 1.0 : v;
  return v > kMaxScaling ? kMaxScaling : v;
}

// ---------------------------------------------------------------------------------------
// sparse kernels over one instance.  A rows go through the CSR view (values fetched
// through the CSR->CSC permutation), A and P columns through the CSC arrays.
// ---------------------------------------------------------------------------------------
template <typename F>
__device__ __forceinline__ void for_rows_A(const PatternDev& P, const double* __restrict__ Aval,
                                           const double* __restrict__ src, F f) {
  const int tid = threadIdx.x, T = blockDim.x;
  // long rows: 4 lanes per row
  for (int base = 0; base < P.n_long; base += (T >> 2)) {
    const int idx = base + (tid >> 2);
    const bool valid = idx < P.n_long;
    double s = 0.0;
    int row = 0;
    if (valid) {
      row = P.rows_long[idx];
      const int e = P.a_rowptr[row + 1];
#pragma unroll 4
      for (int k = P.a_rowptr[row] + (tid & 3); k < e; k += 4) s += Aval[P.a_perm[k]] * src[P.a_colidx[k]];
    }
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    if (valid && (tid & 3) == 0) f(row, s);
  }
  for (int idx = tid; idx < P.n_short; idx += T) {
    const int row = P.rows_short[idx];
    double s = 0.0;
    const int e = P.a_rowptr[row + 1];
#pragma unroll 4
    for (int k = P.a_rowptr[row]; k < e; ++k) s += Aval[P.a_perm[k]] * src[P.a_colidx[k]];
    f(row, s);
  }
}

__device__ __forceinline__ double col_dot_A(const PatternDev& P, const double* __restrict__ Aval,
                                            const double* __restrict__ v, int j) {
  double s = 0.0;
  const int e = P.a_colptr[j + 1];
#pragma unroll 4
  for (int k = P.a_colptr[j]; k < e; ++k) s += Aval[k] * v[P.a_rowidx[k]];
  return s;
}
__device__ __forceinline__ double col_dot_P(const PatternDev& P, const double* __restrict__ Pval,
                                            const double* __restrict__ v, int j) {
  double s = 0.0;
  const int e = P.p_colptr[j + 1];
#pragma unroll 4
  for (int k = P.p_colptr[j]; k < e; ++k) s += Pval[k] * v[P.p_rowidx[k]];
  return s;
}

// cycle accounting of CTA 0 (diagnostics; see ocp_b200_get_phase_cycles)
struct PhaseClock {
  long long* out; long long t0;
  __device__ PhaseClock(long long* o) : out((blockIdx.x == 0 && threadIdx.x == 0) ? o : nullptr), t0(0) { if (out) t0 = clock64(); }
  __device__ __forceinline__ void lap(int phase) {
    if (out) { const long long t = clock64(); out[phase] += t - t0; t0 = t; }
  }
};


// The finer split of the solve / factor phases costs a predicated branch per lap in every thread;
// it is compiled in only with -DOCP_B200_FINE_PHASES (tools/phase_profile.py documents the numbers).
#ifdef OCP_B200_FINE_PHASES
#define OCP_B200_FINE_CLOCK(name, ptr) PhaseClock name(ptr)
#define OCP_B200_FINE_LAP(name, phase) name.lap(phase)
#else
#define OCP_B200_FINE_CLOCK(name, ptr)
#define OCP_B200_FINE_LAP(name, phase)
#endif

struct QpResult {
  int status, iters, pcg_iters, rho_updates, checks;
  double prim_res, dual_res, rho;
};

// per-instance outputs after a QP: solution, SQP update x += alpha * d[np:], info / stats.
// xs / ys: unscaled primal / dual solution of the instance (any address space).
__device__ inline void write_outputs(const PatternDev& P, const SolveArgs& A, const double* xs, const double* ys,
                                      Reducer& R, int inst, const QpResult& res) {
  const bool solved_setup = res.status != OCP_B200_QP_UNSOLVED || res.iters > 0;
  if (A.sol_x)
    for (int j = threadIdx.x; j < P.n; j += blockDim.x) A.sol_x[size_t(inst) * P.n + j] = solved_setup ? xs[j] : 0.0;
  if (A.sol_y)
    for (int i = threadIdx.x; i < P.m; i += blockDim.x) A.sol_y[size_t(inst) * P.m + i] = solved_setup ? ys[i] : 0.0;
  double nrm[1] = {0.0};
  if (A.x_iter) {
    double* xi = A.x_iter + size_t(inst) * A.N;
    for (int i = threadIdx.x; i < A.N; i += blockDim.x) {
      const double dx = solved_setup ? A.sqp_alpha * xs[A.np + i] : 0.0;
      xi[i] += dx;
      nrm[0] += dx * dx;
    }
    block_reduce<1, false>(nrm, R);
  }
  if (threadIdx.x == 0) {
    if (A.info) {
      double* f = A.info + size_t(inst) * OCP_B200_NINFO;
      f[OCP_B200_INFO_STATUS] = res.status; f[OCP_B200_INFO_ITERS] = res.iters;
      f[OCP_B200_INFO_PCG_ITERS] = res.pcg_iters; f[OCP_B200_INFO_PRIM_RES] = res.prim_res;
      f[OCP_B200_INFO_DUAL_RES] = res.dual_res; f[OCP_B200_INFO_RHO] = res.rho;
      f[OCP_B200_INFO_RHO_UPDATES] = res.rho_updates; f[OCP_B200_INFO_CHECKS] = res.checks;
    }
    if (A.stats) {
      double* s = A.stats + size_t(inst) * OCP_B200_NSTATS;
      if (A.first_step) for (int k = 0; k < OCP_B200_NSTATS; ++k) s[k] = 0.0;
      s[OCP_B200_STAT_QP_STATUS] = res.status;
      s[OCP_B200_STAT_SQP_STEPS] += 1.0;
      s[OCP_B200_STAT_ADMM_ITERS] += res.iters;
      s[OCP_B200_STAT_PCG_ITERS] += res.pcg_iters;
      s[OCP_B200_STAT_PRIM_RES] = res.prim_res;
      s[OCP_B200_STAT_DUAL_RES] = res.dual_res;
      s[OCP_B200_STAT_RHO_UPDATES] += res.rho_updates;
      s[OCP_B200_STAT_LAST_ADMM] = res.iters;
      s[OCP_B200_STAT_LAST_RHO] = res.rho;
      s[OCP_B200_STAT_CHECKS] += res.checks;
      s[OCP_B200_STAT_STEP_NORM] = sqrt(nrm[0]);
    }
  }
  __syncthreads();
}

}  // namespace ocpb200
