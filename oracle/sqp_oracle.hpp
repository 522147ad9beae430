// ORACLE -- TEST INFRASTRUCTURE ONLY (see osqp_restate.hpp for the rule).
//
// CPU restatement of the reference's SQP driver and QP adapter:
//   * constructor  = SQPOptimizationSolver::SQPOptimizationSolver
//                    (src/sqp_solver/SQPOptimizationSolver.cpp:12-92) with the AutoDifferentiator
//                    calls written out (src/sqp_solver/AutoDifferentiator.cpp:16-27, 132-140):
//                    w = [p; x], c = [p; x; g], H = hess_w f, grad, J = dc/dw, l' = l + (-c);
//   * local_system = getLocalSystem (:100-120), evaluated by casadi-lite's tape interpreter
//                    in place of CasADi's SX virtual machine;
//   * solve        = getOptimalSolution (:127-216) with CuCaQP::setSystem/initSolver/solve
//                    (src/sqp_solver/CuCaQP.cpp:271-288, 183-224) inlined: values cast to Real
//                    (CuCaQP.h:128, 149), a brand-new cold-started OSQP set-up every step,
//                    x += alpha * d[np:], exactly step_num steps, QP status ignored.
// CasADi / OSQP are absent from this image: parity against them is UNPINNED; the anchors are
// the reference's own test/test.cpp cases (tests/test_oracle_kat.py).
#pragma once

#include <casadi/casadi.hpp>

#include "osqp_restate.hpp"

namespace oracle {

struct SqpStats {
  int qp_status = 0, sqp_steps = 0, last_admm = 0;
  long admm_iters = 0, rho_updates = 0, checks = 0;
  double prim_res = 0, dual_res = 0, objective = 0, last_rho = 0, step_norm = 0;
};

template <typename Real>
class SqpReference {
 public:
  SqpReference(casadi::SXDict& nlp, int step_num, double alpha) : stepNum_(step_num), alpha_(alpha) {
    using casadi::SX;
    if (nlp.find("f") == nlp.end()) throw std::invalid_argument("objective 'f' is not defined");
    if (nlp.find("x") == nlp.end()) throw std::invalid_argument("decision variables 'x' are not defined");
    SX f = nlp["f"], x = nlp["x"];
    SX g = nlp.find("g") != nlp.end() ? nlp["g"] : SX();
    SX p = nlp.find("p") != nlp.end() ? nlp["p"] : SX();
    objective_ = casadi::Function("objective", {p, x}, {f});
    SX w = SX::vertcat({p, x});
    SX c = SX::vertcat({p, x, g});
    SX grad;
    SX H = SX::hessian(f, w, grad);
    SX J = SX::jacobian(c, w);
    SX b = -c;
    SX l = SX::sym("l", c.size1()), u = SX::sym("u", c.size1());
    local_ = casadi::Function("localSystemFunction", {p, x, l, u}, {H, grad, J, l + b, u + b});
    np_ = static_cast<int>(p.numel()); N_ = static_cast<int>(x.numel()); ng_ = static_cast<int>(g.numel());
    n_ = np_ + N_; m_ = n_ + ng_;
    qp_settings.eps_abs = 1e-3; qp_settings.eps_rel = 1e-3; qp_settings.max_iter = 10000;  // :83-85
    const casadi::Sparsity& hs = local_.sparsity_out(0);
    const casadi::Sparsity& as = local_.sparsity_out(2);
    hp_.assign(hs.get_colind().begin(), hs.get_colind().end()); hi_.assign(hs.get_row().begin(), hs.get_row().end());
    ap_.assign(as.get_colind().begin(), as.get_colind().end()); ai_.assign(as.get_row().begin(), as.get_row().end());
  }

  int np() const { return np_; }
  int N() const { return N_; }
  int ng() const { return ng_; }
  int n() const { return n_; }
  int m() const { return m_; }
  const std::vector<int>& h_colptr() const { return hp_; }
  const std::vector<int>& h_rowidx() const { return hi_; }
  const std::vector<int>& a_colptr() const { return ap_; }
  const std::vector<int>& a_rowidx() const { return ai_; }
  const casadi::Function& localSystemFunction() const { return local_; }
  int stepNum() const { return stepNum_; }
  double alpha() const { return alpha_; }
  void setSchedule(int step_num, double alpha) { stepNum_ = step_num; alpha_ = alpha; }

  OsqpSettings qp_settings;

  // per-thread scratch
  struct Work {
    std::vector<double> w, lfull, ufull, hv, q, av, l, u;
    Csc<Real> P, A;
    std::vector<Real> qr, lr, ur, d, y;
    OsqpRestated<Real> qp;
    std::vector<TraceRecord> trace;
  };
  void init_work(Work& k) const {
    k.w.resize(local_.sz_w()); k.lfull.resize(m_); k.ufull.resize(m_);
    k.hv.resize(hi_.size()); k.q.resize(n_); k.av.resize(ai_.size()); k.l.resize(m_); k.u.resize(m_);
    k.P.nrow = k.P.ncol = n_; k.P.p = hp_; k.P.i = hi_; k.P.x.resize(hi_.size());
    k.A.nrow = m_; k.A.ncol = n_; k.A.p = ap_; k.A.i = ai_; k.A.x.resize(ai_.size());
    k.qr.resize(n_); k.lr.resize(m_); k.ur.resize(m_); k.d.resize(n_); k.y.resize(m_);
  }

  // getLocalSystem: l = [p; lbx; lbg], u = [p; ubx; ubg]
  void local_system(Work& k, const double* p, const double* x, const double* lbx, const double* ubx,
                    const double* lbg, const double* ubg) const {
    for (int i = 0; i < np_; ++i) { k.lfull[i] = p[i]; k.ufull[i] = p[i]; }
    for (int i = 0; i < N_; ++i) { k.lfull[np_ + i] = lbx[i]; k.ufull[np_ + i] = ubx[i]; }
    for (int i = 0; i < ng_; ++i) { k.lfull[n_ + i] = lbg[i]; k.ufull[n_ + i] = ubg[i]; }
    const double* arg[4] = {p, x, k.lfull.data(), k.ufull.data()};
    double* res[5] = {k.hv.data(), k.q.data(), k.av.data(), k.l.data(), k.u.data()};
    local_.eval(arg, res, k.w.data());
  }

  double objective(Work& k, const double* p, const double* x) const {
    std::vector<double> w(objective_.sz_w());
    const double* arg[2] = {p, x};
    double f = 0;
    double* res[1] = {&f};
    objective_.eval(arg, res, w.data());
    (void)k;
    return f;
  }

  // getOptimalSolution on raw buffers; x_inout plays the role of the persistent result_["x"]
  void solve(Work& k, const double* p, const double* lbx, const double* ubx, const double* lbg,
             const double* ubg, double* x_inout, double* f_out, SqpStats* stats, bool keep_trace = false) const {
    SqpStats st;
    k.qp.settings = qp_settings;
    for (int step = 0; step < stepNum_; ++step) {
      local_system(k, p, x_inout, lbx, ubx, lbg, ubg);
      // CuCaQP::setSystem: CCS order kept, values cast to OSQPFloat
      for (size_t e = 0; e < k.hv.size(); ++e) k.P.x[e] = static_cast<Real>(k.hv[e]);
      for (size_t e = 0; e < k.av.size(); ++e) k.A.x[e] = static_cast<Real>(k.av[e]);
      for (int j = 0; j < n_; ++j) k.qr[j] = static_cast<Real>(k.q[j]);
      for (int i = 0; i < m_; ++i) { k.lr[i] = static_cast<Real>(k.l[i]); k.ur[i] = static_cast<Real>(k.u[i]); }
      k.trace.clear();
      k.qp.trace = keep_trace ? &k.trace : nullptr;
      bool ok = k.qp.setup(k.P, k.qr.data(), k.A, k.lr.data(), k.ur.data());   // initSolver()
      if (ok) {
        k.qp.solve(k.d.data(), k.y.data());                                     // solve()
      } else {
        // OsqpEigen leaves its solution vector untouched when set-up fails; the reference then
        // adds whatever it holds.  A zero step is the closest deterministic equivalent.
        std::fill(k.d.begin(), k.d.end(), Real(0));
        k.qp.info = OsqpInfo();
      }
      double nrm = 0;
      for (int i = 0; i < N_; ++i) {
        double dx = alpha_ * static_cast<double>(k.d[np_ + i]);  // result_["x"] += alpha * solution[pSize:]
        x_inout[i] += dx;
        nrm += dx * dx;
      }
      st.step_norm = std::sqrt(nrm);
      st.qp_status = k.qp.info.status; st.last_admm = k.qp.info.iter;
      st.admm_iters += k.qp.info.iter; st.rho_updates += k.qp.info.rho_updates; st.checks += k.qp.info.checks;
      st.prim_res = k.qp.info.prim_res; st.dual_res = k.qp.info.dual_res; st.last_rho = k.qp.info.rho;
      st.sqp_steps = step + 1;
    }
    st.objective = objective(k, p, x_inout);
    if (f_out) *f_out = st.objective;
    if (stats) *stats = st;
  }

 private:
  int stepNum_;
  double alpha_;
  int np_ = 0, N_ = 0, ng_ = 0, n_ = 0, m_ = 0;
  casadi::Function objective_, local_;
  std::vector<int> hp_, hi_, ap_, ai_;
};

}  // namespace oracle
