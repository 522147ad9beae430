// ORACLE -- TEST INFRASTRUCTURE ONLY.  Nothing in the product path (optimal_control_problem_b200/,
// include/) may include, link or call this file; only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs use it, as the checker and the CPU baseline.
//
// CPU restatement of the ADMM QP solver the reference reaches through OsqpEigen
// (src/sqp_solver/CuCaQP.cpp:183-224: initSolver -> osqp_setup, solve -> osqp_solve):
// OSQP v1.0.0.beta1, `builtin` algebra = direct LDL' solve of the quasi-definite KKT system
// (cpu_install.sh:34-44).  OSQP's source is NOT under /root/reference and not installed in
// this image, so this is a restatement of the published algorithm
// (Stellato et al., "OSQP: an operator splitting solver for quadratic programs", 2020, and
// the v0.6/v1.0 sources as remembered), not a compilation of it:
//
//   PARITY UNPINNED against OSQP itself.  What pins this file: the 7 analytic optima of the
//   reference's own test/test.cpp:13-185 (tests/test_oracle_kat.py), an independent
//   active-set / dense-KKT cross-check in numpy (tests/test_oracle_qp.py), and the
//   KKT-residual property tests.  Choices that upstream leaves to wall-clock time are made
//   deterministic and documented where they occur.  tools/pin_reference.py produces OSQP /
//   CasADi fixtures (tests/golden/ref_*.npz) on a machine that has the libraries;
//   tests/test_reference_pins.py compares this file with them when they exist.
//
// Algorithm:
//   setup : clamp l,u to +-1e30; Ruiz equilibration x`scaling` with cost normalisation;
//           rho vector by constraint type; KKT = [[P+sigma I, A'],[A, -diag(1/rho)]] ordered
//           by approximate minimum degree (AMD), LDL' by the QDLDL up-looking algorithm.
//   iterate: rhs = (sigma x - q, z - y/rho); solve; z~ = z + (nu - y)/rho;
//           x+ = a x~ + (1-a) x;  z+ = clip(a z~ + (1-a) z + y/rho, l, u);
//           y+ = y + rho (a z~ + (1-a) z - z+)                       (a = 1.6)
//   every `check_termination` iterations: unscaled residuals, termination test,
//           infeasibility certificates; every `adaptive_rho_interval`: rho re-estimate,
//           applied (with a numeric refactorisation) when it moved by more than 5x.
// The template parameter mirrors OSQP_USE_FLOAT (cpu_install.sh:44): Real = float is the
// reference's build, Real = double is the parity oracle.
#pragma once

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <limits>
#include <vector>

namespace oracle {

struct OsqpSettings {
  double rho = 0.1, sigma = 1e-6, alpha = 1.6;
  double eps_abs = 1e-3, eps_rel = 1e-3, eps_prim_inf = 1e-4, eps_dual_inf = 1e-4;
  int max_iter = 10000, scaling = 10, check_termination = 25;
  int adaptive_rho = 1;
  // Upstream (adaptive_rho_interval = 0, profiling on, cuda_install.sh:40) picks the interval
  // from measured wall-clock time, so the reference itself is not reproducible run to run.
  // 0 here selects upstream's deterministic fallback: ADAPTIVE_RHO_MULTIPLE_TERMINATION (4)
  // x check_termination = 100.
  int adaptive_rho_interval = 0;
  double adaptive_rho_tolerance = 5.0;
  bool reuse_symbolic = false;  // keep ordering + elimination tree between setups (baseline variant)
};

enum OsqpStatus {
  OSQP_SOLVED = 1, OSQP_SOLVED_INACCURATE = 2, OSQP_PRIMAL_INFEASIBLE = 3, OSQP_DUAL_INFEASIBLE = 5,
  OSQP_MAX_ITER_REACHED = 7, OSQP_UNSOLVED = 11
};

struct OsqpInfo {
  int status = OSQP_UNSOLVED, iter = 0, rho_updates = 0, checks = 0;
  double prim_res = 0, dual_res = 0, rho = 0, obj_val = 0;
};

struct TraceRecord { int iter; double prim_res, dual_res, rho; int status; };

template <typename Real>
struct Csc {
  int nrow = 0, ncol = 0;
  std::vector<int> p, i;
  std::vector<Real> x;
};

// ---- QDLDL (restated): elimination tree, up-looking LDL', triangular solves ---------------
template <typename Real>
struct Ldl {
  int n = 0;
  std::vector<int> etree, Lnz, Lp, Li;
  std::vector<Real> Lx, D, Dinv;
  std::vector<int> iwork;
  std::vector<unsigned char> bwork;
  std::vector<Real> fwork;

  // A: upper triangular CSC with a full diagonal
  bool symbolic(const Csc<Real>& A) {
    n = A.ncol;
    etree.assign(n, -1); Lnz.assign(n, 0);
    std::vector<int> work(n, 0);
    for (int j = 0; j < n; ++j) {
      work[j] = j;
      if (A.p[j] == A.p[j + 1]) return false;
      for (int q = A.p[j]; q < A.p[j + 1]; ++q) {
        int i = A.i[q];
        if (i > j) return false;
        while (work[i] != j) {
          if (etree[i] == -1) etree[i] = j;
          Lnz[i]++;
          work[i] = j;
          i = etree[i];
        }
      }
    }
    Lp.assign(n + 1, 0);
    for (int i = 0; i < n; ++i) Lp[i + 1] = Lp[i] + Lnz[i];
    Li.assign(Lp[n], 0); Lx.assign(Lp[n], Real(0));
    D.assign(n, Real(0)); Dinv.assign(n, Real(0));
    iwork.assign(3 * n, 0); bwork.assign(n, 0); fwork.assign(n, Real(0));
    return true;
  }

  bool numeric(const Csc<Real>& A) {
    int* yIdx = iwork.data();
    int* elim = iwork.data() + n;
    int* next = iwork.data() + 2 * n;
    Real* y = fwork.data();
    for (int i = 0; i < n; ++i) { bwork[i] = 0; y[i] = Real(0); D[i] = Real(0); next[i] = Lp[i]; }
    for (int k = 0; k < n; ++k) {
      int nnzY = 0;
      for (int q = A.p[k]; q < A.p[k + 1]; ++q) {
        int b = A.i[q];
        if (b == k) { D[k] = A.x[q]; continue; }
        y[b] = A.x[q];
        int nx = b;
        if (!bwork[nx]) {
          bwork[nx] = 1;
          elim[0] = nx;
          int nnzE = 1;
          nx = etree[b];
          while (nx != -1 && nx < k) {
            if (bwork[nx]) break;
            bwork[nx] = 1;
            elim[nnzE++] = nx;
            nx = etree[nx];
          }
          while (nnzE) yIdx[nnzY++] = elim[--nnzE];
        }
      }
      for (int t = nnzY - 1; t >= 0; --t) {
        int c = yIdx[t];
        int tmp = next[c];
        Real yc = y[c];
        for (int j = Lp[c]; j < tmp; ++j) y[Li[j]] -= Lx[j] * yc;
        Li[tmp] = k;
        Lx[tmp] = yc * Dinv[c];
        D[k] -= yc * Lx[tmp];
        next[c]++;
        y[c] = Real(0);
        bwork[c] = 0;
      }
      if (D[k] == Real(0)) return false;
      Dinv[k] = Real(1) / D[k];
    }
    return true;
  }

  void solve(Real* x) const {
    for (int i = 0; i < n; ++i) {
      Real xi = x[i];
      for (int j = Lp[i]; j < Lp[i + 1]; ++j) x[Li[j]] -= Lx[j] * xi;
    }
    for (int i = 0; i < n; ++i) x[i] *= Dinv[i];
    for (int i = n - 1; i >= 0; --i) {
      Real xi = x[i];
      for (int j = Lp[i]; j < Lp[i + 1]; ++j) xi -= Lx[j] * x[Li[j]];
      x[i] = xi;
    }
  }
};

// Fill-reducing ordering of a symmetric matrix given by its upper triangle: approximate minimum degree on
// the quotient graph (Amestoy, Davis, Duff, "An approximate minimum degree ordering algorithm", SIMAX 1996 --
// the algorithm behind SuiteSparse AMD, which upstream OSQP calls through QDLDL's interface).  Eliminated
// variables become elements; a variable's degree is bounded by
//     d_i = min(n - k,  d_i + |L_p \ i|,  |A_i \ i| + |L_p \ i| + sum_{e in E_i \ p} |L_e \ L_p|)
// with |L_e \ L_p| obtained for all elements at once (the w(e) pass); elements whose variables all lie in
// L_p are absorbed.  No supervariables / mass elimination: the KKT systems here have ~10^3 rows.  Any
// fill-reducing ordering leaves the ADMM iterates unchanged up to rounding; what this buys over the exact
// greedy minimum degree that round 1 used is set-up TIME (the CPU baseline re-runs it every SQP step, as the
// reference does: CuCaQP.cpp:273-276).
inline std::vector<int> min_degree_order(int n, const std::vector<int>& Ap, const std::vector<int>& Ai) {
  // the lists keep their capacity from call to call (one set per thread): the ordering is re-run for every QP
  struct Lists { std::vector<std::vector<int>> A, E, L, bucket; };
  static thread_local Lists ws;
  auto reset = [n](std::vector<std::vector<int>>& v, size_t count) {
    if (v.size() < count) v.resize(count);
    for (size_t i = 0; i < count; ++i) v[i].clear();
    (void)n;
  };
  reset(ws.A, n); reset(ws.E, n); reset(ws.L, n); reset(ws.bucket, size_t(n) + 1);
  std::vector<std::vector<int>>&A = ws.A, &E = ws.E, &L = ws.L, &bucket = ws.bucket;
  for (int j = 0; j < n; ++j)
    for (int q = Ap[j]; q < Ap[j + 1]; ++q) {
      const int i = Ai[q];
      if (i != j) { A[i].push_back(j); A[j].push_back(i); }
    }
  std::vector<int> deg(n), stamp(n, -1), wst(n, -1), w(n, 0), perm;
  std::vector<char> elim(n, 0), absorbed(n, 0);
  for (int v = 0; v < n; ++v) {
    std::sort(A[v].begin(), A[v].end());
    A[v].erase(std::unique(A[v].begin(), A[v].end()), A[v].end());
    deg[v] = static_cast<int>(A[v].size());
    bucket[deg[v]].push_back(v);
  }
  perm.reserve(n);
  std::vector<int> Lp;
  int mindeg = 0;
  for (int k = 0; k < n; ++k) {
    int p = -1;
    while (p < 0) {
      while (mindeg <= n && bucket[mindeg].empty()) ++mindeg;
      const int c = bucket[mindeg].back();
      bucket[mindeg].pop_back();
      if (!elim[c] && deg[c] == mindeg) p = c;   // stale entries (degree changed, already eliminated) are skipped
    }
    // L_p = (A_p  u  union of L_e, e in E_p) \ p ; the elements of E_p are absorbed into p
    stamp[p] = k;
    Lp.clear();
    for (int i : A[p]) if (!elim[i] && stamp[i] != k) { stamp[i] = k; Lp.push_back(i); }
    for (int e : E[p]) {
      if (absorbed[e]) continue;
      for (int i : L[e]) if (!elim[i] && stamp[i] != k) { stamp[i] = k; Lp.push_back(i); }
      absorbed[e] = 1;
      L[e].clear();
    }
    elim[p] = 1;
    perm.push_back(p);
    A[p].clear();
    E[p].clear();
    // w(e) = |L_e \ L_p| for every element adjacent to a variable of L_p
    for (int i : Lp)
      for (int e : E[i]) {
        if (absorbed[e]) continue;
        if (wst[e] != k) { wst[e] = k; w[e] = static_cast<int>(L[e].size()); }
        --w[e];
      }
    const int lp = static_cast<int>(Lp.size());
    for (int i : Lp) {
      // A_i loses what L_p now covers; E_i loses absorbed elements (aggressively: those inside L_p) and gains p
      size_t o = 0;
      for (int j : A[i]) if (!elim[j] && stamp[j] != k) A[i][o++] = j;
      A[i].resize(o);
      long d = static_cast<long>(o) + (lp - 1);
      o = 0;
      for (int e : E[i]) {
        if (absorbed[e]) continue;
        if (w[e] == 0) { absorbed[e] = 1; L[e].clear(); continue; }
        E[i][o++] = e;
        d += w[e];
      }
      E[i].resize(o);
      E[i].push_back(p);
      d = std::min<long>(d, std::min<long>(n - k - 1, static_cast<long>(deg[i]) + lp - 1));
      deg[i] = static_cast<int>(std::max<long>(d, 0));
      bucket[deg[i]].push_back(i);
      if (deg[i] < mindeg) mindeg = deg[i];
    }
    L[p] = Lp;
  }
  return perm;
}

template <typename Real>
class OsqpRestated {
 public:
  static constexpr double kInfty = 1e30, kMinScaling = 1e-4, kMaxScaling = 1e4;
  static constexpr double kRhoMin = 1e-6, kRhoMax = 1e6, kRhoTol = 1e-4, kRhoEqOverIneq = 1e3;
  static constexpr double kDivisionTol = 1e-30;  // 1/OSQP_INFTY

  OsqpSettings settings;
  OsqpInfo info;
  std::vector<TraceRecord>* trace = nullptr;

  // P: n-by-n, only its UPPER triangle is read (OsqpEigen hands OSQP the upper triangle);
  // A: m-by-n.  Copies everything, like osqp_setup.
  bool setup(const Csc<Real>& Pin, const Real* q_in, const Csc<Real>& Ain, const Real* l_in, const Real* u_in) {
    n_ = Ain.ncol; m_ = Ain.nrow;
    // upper triangle of P
    P_.nrow = P_.ncol = n_; P_.p.assign(n_ + 1, 0); P_.i.clear(); P_.x.clear();
    for (int j = 0; j < n_; ++j) {
      for (int k = Pin.p[j]; k < Pin.p[j + 1]; ++k)
        if (Pin.i[k] <= j) { P_.i.push_back(Pin.i[k]); P_.x.push_back(Pin.x[k]); }
      P_.p[j + 1] = static_cast<int>(P_.i.size());
    }
    A_ = Ain;
    q_.assign(q_in, q_in + n_);
    l_.assign(l_in, l_in + m_); u_.assign(u_in, u_in + m_);
    for (int i = 0; i < m_; ++i) {
      if (l_[i] > u_[i]) return false;  // validate_data: lower bound above upper bound
      l_[i] = std::max<Real>(l_[i], Real(-kInfty));
      u_[i] = std::min<Real>(u_[i], Real(kInfty));
    }
    scale();
    rho_ = std::min<Real>(std::max<Real>(Real(settings.rho), Real(kRhoMin)), Real(kRhoMax));
    ctype_.assign(m_, 0); rho_vec_.assign(m_, 0); rho_inv_.assign(m_, 0);
    for (int i = 0; i < m_; ++i) {
      if (l_[i] < -kInfty * kMinScaling && u_[i] > kInfty * kMinScaling) ctype_[i] = -1;
      else if (u_[i] - l_[i] < kRhoTol) ctype_[i] = 1;
      else ctype_[i] = 0;
    }
    set_rho_vec();
    if (!build_kkt()) return false;
    x_.assign(n_, 0); z_.assign(m_, 0); y_.assign(m_, 0);       // cold start
    xprev_.assign(n_, 0); zprev_.assign(m_, 0);
    xz_.assign(n_ + m_, 0); dx_.assign(n_, 0); dy_.assign(m_, 0);
    Ax_.assign(m_, 0); Px_.assign(n_, 0); Aty_.assign(n_, 0);
    tmp_n_.assign(n_, 0); tmp_m_.assign(m_, 0); sol_.assign(n_ + m_, 0);
    info = OsqpInfo();
    info.rho = rho_;
    return true;
  }

  void solve(Real* x_out, Real* y_out) {
    const OsqpSettings& s = settings;
    const Real alpha = Real(s.alpha), sigma = Real(s.sigma);
    int rho_interval = s.adaptive_rho_interval > 0 ? s.adaptive_rho_interval
                                                    : (s.check_termination > 0 ? 4 * s.check_termination : 100);   // ADAPTIVE_RHO_FIXED
    int iter = 0;
    bool exited = false;
    for (iter = 1; iter <= s.max_iter; ++iter) {
      xprev_.swap(x_); zprev_.swap(z_);
      // x~, nu from the KKT system
      for (int j = 0; j < n_; ++j) xz_[j] = sigma * xprev_[j] - q_[j];
      for (int i = 0; i < m_; ++i) xz_[n_ + i] = zprev_[i] - rho_inv_[i] * y_[i];
      kkt_solve(xz_.data());
      for (int i = 0; i < m_; ++i) xz_[n_ + i] = zprev_[i] + rho_inv_[i] * (xz_[n_ + i] - y_[i]);  // z~
      for (int j = 0; j < n_; ++j) {
        x_[j] = alpha * xz_[j] + (Real(1) - alpha) * xprev_[j];
        dx_[j] = x_[j] - xprev_[j];
      }
      for (int i = 0; i < m_; ++i) {
        Real zr = alpha * xz_[n_ + i] + (Real(1) - alpha) * zprev_[i];
        Real v = zr + rho_inv_[i] * y_[i];
        z_[i] = std::min(std::max(v, l_[i]), u_[i]);
        dy_[i] = rho_vec_[i] * (zr - z_[i]);
        y_[i] += dy_[i];
      }
      const bool can_check = s.check_termination && (iter % s.check_termination == 0);
      bool have_info = false;
      if (can_check) {
        update_info(iter);
        have_info = true;
        if (check_termination(false)) { exited = true; break; }
      }
      if (s.adaptive_rho && rho_interval && (iter % rho_interval == 0)) {
        if (!have_info) update_info(iter);
        adapt_rho();
      }
    }
    if (!exited) {
      iter = s.max_iter;
      if (!(s.check_termination && (iter % s.check_termination == 0))) update_info(iter);
      if (!check_termination(false)) {
        if (!check_termination(true)) info.status = OSQP_MAX_ITER_REACHED;
      }
    }
    info.iter = std::min(iter, s.max_iter);
    info.rho = rho_;
    // store_solution: unscale, or NaN when a certificate was found
    const bool has_sol = info.status != OSQP_PRIMAL_INFEASIBLE && info.status != OSQP_DUAL_INFEASIBLE;
    const Real nanv = std::numeric_limits<Real>::quiet_NaN();
    for (int j = 0; j < n_; ++j) x_out[j] = has_sol ? D_[j] * x_[j] : nanv;
    if (y_out)
      for (int i = 0; i < m_; ++i) y_out[i] = has_sol ? cinv_ * E_[i] * y_[i] : nanv;
  }

  int n() const { return n_; }
  int m() const { return m_; }
  long kkt_nnz_L() const { return static_cast<long>(ldl_.Li.size()); }

 private:
  int n_ = 0, m_ = 0;
  Csc<Real> P_, A_;  // scaled data (P upper triangular)
  std::vector<Real> q_, l_, u_, D_, E_, Dinv_, Einv_;
  Real c_ = 1, cinv_ = 1, rho_ = Real(0.1);
  std::vector<int> ctype_;
  std::vector<Real> rho_vec_, rho_inv_;
  std::vector<Real> x_, z_, y_, xprev_, zprev_, xz_, dx_, dy_, Ax_, Px_, Aty_, tmp_n_, tmp_m_, sol_;
  // KKT
  Csc<Real> K_;                       // permuted upper triangular KKT
  std::vector<int> perm_, iperm_;
  std::vector<int> rho_pos_;          // position in K_.x of the -1/rho_i diagonal entries
  std::vector<int> Pmap_, Amap_, sig_pos_;  // positions of P / A / sigma entries in K_.x
  Ldl<Real> ldl_;
  bool have_symbolic_ = false;

  static Real limit(Real v) {
    v = v < Real(kMinScaling) ? Real(1) : v;
    return v > Real(kMaxScaling) ? Real(kMaxScaling) : v;
  }

  void col_norms_sym_triu(std::vector<Real>& out) const {
    std::fill(out.begin(), out.end(), Real(0));
    for (int j = 0; j < n_; ++j)
      for (int k = P_.p[j]; k < P_.p[j + 1]; ++k) {
        Real a = std::fabs(P_.x[k]);
        int i = P_.i[k];
        out[j] = std::max(out[j], a);
        if (i != j) out[i] = std::max(out[i], a);
      }
  }

  void scale() {
    D_.assign(n_, 1); E_.assign(m_, 1); c_ = 1;
    std::vector<Real> Dt(n_), Et(m_);
    for (int it = 0; it < settings.scaling; ++it) {
      col_norms_sym_triu(Dt);
      std::fill(Et.begin(), Et.end(), Real(0));
      for (int j = 0; j < n_; ++j)
        for (int k = A_.p[j]; k < A_.p[j + 1]; ++k) {
          Real a = std::fabs(A_.x[k]);
          Dt[j] = std::max(Dt[j], a);
          Et[A_.i[k]] = std::max(Et[A_.i[k]], a);
        }
      for (int j = 0; j < n_; ++j) Dt[j] = Real(1) / std::sqrt(limit(Dt[j]));
      for (int i = 0; i < m_; ++i) Et[i] = Real(1) / std::sqrt(limit(Et[i]));
      for (int j = 0; j < n_; ++j) {
        for (int k = P_.p[j]; k < P_.p[j + 1]; ++k) P_.x[k] *= Dt[P_.i[k]] * Dt[j];
        for (int k = A_.p[j]; k < A_.p[j + 1]; ++k) A_.x[k] *= Et[A_.i[k]] * Dt[j];
        q_[j] *= Dt[j];
        D_[j] *= Dt[j];
      }
      for (int i = 0; i < m_; ++i) E_[i] *= Et[i];
      // cost normalisation
      col_norms_sym_triu(Dt);
      Real mean = 0;
      for (int j = 0; j < n_; ++j) mean += Dt[j];
      mean /= Real(n_);
      Real qn = 0;
      for (int j = 0; j < n_; ++j) qn = std::max(qn, std::fabs(q_[j]));
      qn = limit(qn);
      Real ct = Real(1) / limit(std::max(mean, qn));
      for (Real& v : P_.x) v *= ct;
      for (Real& v : q_) v *= ct;
      c_ *= ct;
    }
    cinv_ = Real(1) / c_;
    Dinv_.resize(n_); Einv_.resize(m_);
    for (int j = 0; j < n_; ++j) Dinv_[j] = Real(1) / D_[j];
    for (int i = 0; i < m_; ++i) { Einv_[i] = Real(1) / E_[i]; l_[i] *= E_[i]; u_[i] *= E_[i]; }
  }

  void set_rho_vec() {
    for (int i = 0; i < m_; ++i) {
      rho_vec_[i] = ctype_[i] == -1 ? Real(kRhoMin) : (ctype_[i] == 1 ? Real(kRhoEqOverIneq) * rho_ : rho_);
      rho_inv_[i] = Real(1) / rho_vec_[i];
    }
  }

  // KKT = [[P + sigma I, A'], [A, -diag(1/rho)]], upper triangle, symmetric permutation
  bool build_kkt() {
    const int N = n_ + m_;
    const bool reuse = settings.reuse_symbolic && have_symbolic_ && static_cast<int>(perm_.size()) == N;
    // unpermuted upper triangle in triplet-by-column form
    std::vector<int> Kp(N + 1, 0), Ki;
    std::vector<Real> Kx;
    std::vector<int> kind;  // 0..: index into P_.x (+0), A_.x (+nnzP), sigma-only diag (-1), rho diag (-2-i)
    const int nnzP = static_cast<int>(P_.x.size());
    for (int j = 0; j < n_; ++j) {
      bool diag = false;
      for (int k = P_.p[j]; k < P_.p[j + 1]; ++k) {
        Ki.push_back(P_.i[k]); Kx.push_back(P_.x[k] + (P_.i[k] == j ? Real(settings.sigma) : Real(0)));
        kind.push_back(k);
        if (P_.i[k] == j) diag = true;
      }
      if (!diag) { Ki.push_back(j); Kx.push_back(Real(settings.sigma)); kind.push_back(-1); }
      Kp[j + 1] = static_cast<int>(Ki.size());
    }
    // columns n..n+m-1: rows of A, built from the CSC of A
    std::vector<int> rowcnt(m_ + 1, 0);
    for (int k = 0; k < static_cast<int>(A_.i.size()); ++k) rowcnt[A_.i[k] + 1]++;
    for (int i = 0; i < m_; ++i) rowcnt[i + 1] += rowcnt[i];
    std::vector<int> rcol(A_.i.size()), rpos(A_.i.size()), nxt(rowcnt.begin(), rowcnt.end() - 1);
    for (int j = 0; j < n_; ++j)
      for (int k = A_.p[j]; k < A_.p[j + 1]; ++k) { int t = nxt[A_.i[k]]++; rcol[t] = j; rpos[t] = k; }
    for (int i = 0; i < m_; ++i) {
      for (int t = rowcnt[i]; t < rowcnt[i + 1]; ++t) { Ki.push_back(rcol[t]); Kx.push_back(A_.x[rpos[t]]); kind.push_back(nnzP + rpos[t]); }
      Ki.push_back(n_ + i); Kx.push_back(-rho_inv_[i]); kind.push_back(-2 - i);
      Kp[n_ + i + 1] = static_cast<int>(Ki.size());
    }
    if (!reuse) {
      perm_ = min_degree_order(N, Kp, Ki);
      iperm_.assign(N, 0);
      for (int k = 0; k < N; ++k) iperm_[perm_[k]] = k;
    }
    // symmetric permutation into upper triangular form (csc_symperm)
    std::vector<int> cnt(N + 1, 0);
    for (int j = 0; j < N; ++j)
      for (int q = Kp[j]; q < Kp[j + 1]; ++q) {
        int i2 = iperm_[Ki[q]], j2 = iperm_[j];
        cnt[std::max(i2, j2) + 1]++;
      }
    for (int j = 0; j < N; ++j) cnt[j + 1] += cnt[j];
    K_.nrow = K_.ncol = N; K_.p = cnt; K_.i.assign(Ki.size(), 0); K_.x.assign(Ki.size(), Real(0));
    std::vector<int> pos(cnt.begin(), cnt.end() - 1);
    rho_pos_.assign(m_, -1);
    std::vector<int> order(Ki.size());
    for (int j = 0; j < N; ++j)
      for (int q = Kp[j]; q < Kp[j + 1]; ++q) {
        int i2 = iperm_[Ki[q]], j2 = iperm_[j];
        int t = pos[std::max(i2, j2)]++;
        K_.i[t] = std::min(i2, j2); K_.x[t] = Kx[q];
        if (kind[q] <= -2) rho_pos_[-2 - kind[q]] = t;
      }
    // QDLDL needs sorted rows only for the diagonal-last convention it does not rely on; keep as is
    if (!reuse) {
      if (!ldl_.symbolic(K_)) return false;
      have_symbolic_ = true;
    }
    return ldl_.numeric(K_);
  }

  bool refactor_rho() {
    for (int i = 0; i < m_; ++i) K_.x[rho_pos_[i]] = -rho_inv_[i];
    return ldl_.numeric(K_);
  }

  void kkt_solve(Real* b) {
    const int N = n_ + m_;
    for (int k = 0; k < N; ++k) sol_[k] = b[perm_[k]];
    ldl_.solve(sol_.data());
    for (int k = 0; k < N; ++k) b[perm_[k]] = sol_[k];
  }

  void mat_vec_A(const Real* x, Real* y) const {
    for (int i = 0; i < m_; ++i) y[i] = 0;
    for (int j = 0; j < n_; ++j)
      for (int k = A_.p[j]; k < A_.p[j + 1]; ++k) y[A_.i[k]] += A_.x[k] * x[j];
  }
  void mat_tvec_A(const Real* x, Real* y) const {
    for (int j = 0; j < n_; ++j) {
      Real s = 0;
      for (int k = A_.p[j]; k < A_.p[j + 1]; ++k) s += A_.x[k] * x[A_.i[k]];
      y[j] = s;
    }
  }
  void mat_vec_P(const Real* x, Real* y) const {
    for (int j = 0; j < n_; ++j) y[j] = 0;
    for (int j = 0; j < n_; ++j)
      for (int k = P_.p[j]; k < P_.p[j + 1]; ++k) {
        int i = P_.i[k];
        y[i] += P_.x[k] * x[j];
        if (i != j) y[j] += P_.x[k] * x[i];
      }
  }
  static Real norm_inf(const std::vector<Real>& v) {
    Real r = 0;
    for (Real a : v) r = std::max(r, std::fabs(a));
    return r;
  }
  static Real scaled_norm_inf(const std::vector<Real>& s, const std::vector<Real>& v) {
    Real r = 0;
    for (size_t k = 0; k < v.size(); ++k) r = std::max(r, std::fabs(s[k] * v[k]));
    return r;
  }

  // residuals in the UNSCALED problem (scaled_termination = 0); z_prev / x_prev are reused as
  // the residual vectors exactly like upstream, which compute_rho_estimate relies on
  void update_info(int iter) {
    mat_vec_A(x_.data(), Ax_.data());
    for (int i = 0; i < m_; ++i) zprev_[i] = Ax_[i] - z_[i];
    info.prim_res = scaled_norm_inf(Einv_, zprev_);
    mat_vec_P(x_.data(), Px_.data());
    mat_tvec_A(y_.data(), Aty_.data());
    for (int j = 0; j < n_; ++j) xprev_[j] = q_[j] + Px_[j] + Aty_[j];
    info.dual_res = cinv_ * scaled_norm_inf(Dinv_, xprev_);
    Real obj = 0;
    for (int j = 0; j < n_; ++j) obj += x_[j] * (Real(0.5) * Px_[j] + q_[j]);
    info.obj_val = cinv_ * obj;
    info.iter = iter;
    info.checks++;
  }

  bool check_termination(bool approximate) {
    Real eps_abs = Real(settings.eps_abs), eps_rel = Real(settings.eps_rel);
    Real eps_pinf = Real(settings.eps_prim_inf), eps_dinf = Real(settings.eps_dual_inf);
    if (approximate) { eps_abs *= 10; eps_rel *= 10; eps_pinf *= 10; eps_dinf *= 10; }
    const Real eps_prim = eps_abs + eps_rel * std::max(scaled_norm_inf(Einv_, z_), scaled_norm_inf(Einv_, Ax_));
    const Real eps_dual = eps_abs + eps_rel * cinv_ * std::max(scaled_norm_inf(Dinv_, q_),
                                    std::max(scaled_norm_inf(Dinv_, Aty_), scaled_norm_inf(Dinv_, Px_)));
    bool prim_ok = false, dual_ok = false, prim_inf = false, dual_inf = false;
    if (info.prim_res < eps_prim) prim_ok = true; else prim_inf = is_primal_infeasible(eps_pinf);
    if (info.dual_res < eps_dual) dual_ok = true; else dual_inf = is_dual_infeasible(eps_dinf);
    int status = -1;
    if (prim_ok && dual_ok) status = approximate ? OSQP_SOLVED_INACCURATE : OSQP_SOLVED;
    else if (prim_inf) status = OSQP_PRIMAL_INFEASIBLE;
    else if (dual_inf) status = OSQP_DUAL_INFEASIBLE;
    if (trace && !approximate)
      trace->push_back({info.iter, double(info.prim_res), double(info.dual_res), double(rho_), status < 0 ? OSQP_UNSOLVED : status});
    if (status < 0) return false;
    info.status = status;
    return true;
  }

  bool is_primal_infeasible(Real eps) {
    // project dy on the polar of the recession cone of [l, u]
    for (int i = 0; i < m_; ++i) {
      if (u_[i] > kInfty * kMinScaling) {
        if (l_[i] < -kInfty * kMinScaling) dy_[i] = 0;
        else dy_[i] = std::min<Real>(dy_[i], 0);
      } else if (l_[i] < -kInfty * kMinScaling) {
        dy_[i] = std::max<Real>(dy_[i], 0);
      }
    }
    Real norm_dy = scaled_norm_inf(E_, dy_);
    if (norm_dy > kDivisionTol) {
      Real lhs = 0;
      for (int i = 0; i < m_; ++i) lhs += u_[i] * std::max<Real>(dy_[i], 0) + l_[i] * std::min<Real>(dy_[i], 0);
      if (lhs < -eps * norm_dy) {
        mat_tvec_A(dy_.data(), tmp_n_.data());
        return scaled_norm_inf(Dinv_, tmp_n_) < eps * norm_dy;
      }
    }
    return false;
  }

  bool is_dual_infeasible(Real eps) {
    Real norm_dx = scaled_norm_inf(D_, dx_);
    if (norm_dx > kDivisionTol) {
      Real qdx = 0;
      for (int j = 0; j < n_; ++j) qdx += q_[j] * dx_[j];
      if (qdx < -c_ * eps * norm_dx) {
        mat_vec_P(dx_.data(), tmp_n_.data());
        if (scaled_norm_inf(Dinv_, tmp_n_) < c_ * eps * norm_dx) {
          mat_vec_A(dx_.data(), tmp_m_.data());
          for (int i = 0; i < m_; ++i) {
            Real a = Einv_[i] * tmp_m_[i];
            if ((u_[i] < kInfty * kMinScaling && a > eps * norm_dx) ||
                (l_[i] > -kInfty * kMinScaling && a < -eps * norm_dx)) return false;
          }
          return true;
        }
      }
    }
    return false;
  }

  void adapt_rho() {
    // compute_rho_estimate: SCALED residual vectors left in z_prev / x_prev by update_info
    Real pr = norm_inf(zprev_), dr = norm_inf(xprev_);
    Real pn = std::max(norm_inf(z_), norm_inf(Ax_));
    pr /= (pn + Real(1e-10));
    Real dn = std::max(norm_inf(q_), std::max(norm_inf(Aty_), norm_inf(Px_)));
    dr /= (dn + Real(1e-10));
    Real est = rho_ * std::sqrt(pr / (dr + Real(1e-10)));
    est = std::min<Real>(std::max<Real>(est, Real(kRhoMin)), Real(kRhoMax));
    if (est > rho_ * Real(settings.adaptive_rho_tolerance) || est < rho_ / Real(settings.adaptive_rho_tolerance)) {
      rho_ = est;
      set_rho_vec();
      refactor_rho();
      info.rho_updates++;
    }
  }
};

}  // namespace oracle
