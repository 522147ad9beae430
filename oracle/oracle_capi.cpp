// ORACLE -- TEST INFRASTRUCTURE ONLY.  C entry points for tests/, smoke() and bench.py's CPU
// baseline (loaded with ctypes).  Links the host front-end (problem definitions, casadi-lite)
// and the restated CPU solver; contains no CUDA and is never loaded by the product path.
#include <cstring>
#include <memory>
#include <thread>

#include "problems/problems.h"
#include "sqp_oracle.hpp"

using namespace oracle;

namespace {
thread_local std::string g_error;

struct OracleProblem {
  std::string name;
  std::unique_ptr<OptimalControlProblem> ocp;
  std::unique_ptr<SqpReference<double>> ref64;
  std::unique_ptr<SqpReference<float>> ref32;
  std::vector<double> lbx, ubx, lbg, ubg;
  int nf = 0, horizon = 0;
};

// settings vector layout shared with tests/_oracle.py
void apply_settings(const double* s, OsqpSettings& o) {
  if (!s) return;
  o.rho = s[0]; o.sigma = s[1]; o.alpha = s[2]; o.eps_abs = s[3]; o.eps_rel = s[4];
  o.eps_prim_inf = s[5]; o.eps_dual_inf = s[6]; o.max_iter = static_cast<int>(s[7]);
  o.scaling = static_cast<int>(s[8]); o.check_termination = static_cast<int>(s[9]);
  o.adaptive_rho = static_cast<int>(s[10]); o.adaptive_rho_interval = static_cast<int>(s[11]);
  o.adaptive_rho_tolerance = s[12];
}

template <typename Real>
int qp_solve_impl(int n, int m, const int* hp, const int* hi, const double* hx, const double* q, const int* ap,
                  const int* ai, const double* ax, const double* l, const double* u, const double* settings,
                  double* x_out, double* y_out, double* info_out, double* trace_out, int max_trace, int* n_trace) {
  Csc<Real> P, A;
  P.nrow = P.ncol = n; P.p.assign(hp, hp + n + 1); P.i.assign(hi, hi + hp[n]); P.x.resize(hp[n]);
  A.nrow = m; A.ncol = n; A.p.assign(ap, ap + n + 1); A.i.assign(ai, ai + ap[n]); A.x.resize(ap[n]);
  for (int k = 0; k < hp[n]; ++k) P.x[k] = static_cast<Real>(hx[k]);
  for (int k = 0; k < ap[n]; ++k) A.x[k] = static_cast<Real>(ax[k]);
  std::vector<Real> qr(q, q + n), lr(l, l + m), ur(u, u + m), x(n), y(m);
  OsqpRestated<Real> qp;
  apply_settings(settings, qp.settings);
  std::vector<TraceRecord> trace;
  qp.trace = &trace;
  if (!qp.setup(P, qr.data(), A, lr.data(), ur.data())) { g_error = "osqp setup failed (l > u or singular KKT)"; return 1; }
  qp.solve(x.data(), y.data());
  for (int j = 0; j < n; ++j) x_out[j] = x[j];
  if (y_out) for (int i = 0; i < m; ++i) y_out[i] = y[i];
  if (info_out) {
    info_out[0] = qp.info.status; info_out[1] = qp.info.iter; info_out[2] = 0; info_out[3] = qp.info.prim_res;
    info_out[4] = qp.info.dual_res; info_out[5] = qp.info.rho; info_out[6] = qp.info.rho_updates; info_out[7] = qp.info.checks;
  }
  int nt = 0;
  if (trace_out)
    for (const TraceRecord& r : trace) {
      if (nt >= max_trace) break;
      double* t = trace_out + 6 * nt++;
      t[0] = r.iter; t[1] = r.prim_res; t[2] = r.dual_res; t[3] = r.rho; t[4] = 0; t[5] = r.status;
    }
  if (n_trace) *n_trace = nt;
  return 0;
}

void write_stats(const SqpStats& st, double* s) {
  // same layout as OCP_B200_STAT_* in include/ocp_b200.h
  s[0] = st.qp_status; s[1] = st.sqp_steps; s[2] = double(st.admm_iters); s[3] = 0; s[4] = st.prim_res;
  s[5] = st.dual_res; s[6] = st.objective; s[7] = double(st.rho_updates); s[8] = st.last_admm; s[9] = st.last_rho;
  s[10] = double(st.checks); s[11] = st.step_norm;
}

template <typename Real>
void sqp_batch_impl(const OracleProblem& op, const SqpReference<Real>& ref, int B, const double* frames,
                    const double* p, double* x_inout, double* f_out, double* stats, int nthreads) {
  const int N = ref.N(), np = ref.np(), nf = op.nf;
  nthreads = std::max(1, std::min(nthreads, B));
  auto worker = [&](int tid) {
    typename SqpReference<Real>::Work k;
    ref.init_work(k);
    std::vector<double> lbx(op.lbx), ubx(op.ubx);
    for (int b = tid; b < B; b += nthreads) {
      if (frames)
        for (int i = 0; i < nf; ++i) { lbx[i] = frames[size_t(b) * nf + i]; ubx[i] = lbx[i]; }
      SqpStats st;
      double f = 0;
      ref.solve(k, p + size_t(b) * np, lbx.data(), ubx.data(), op.lbg.data(), op.ubg.data(),
                x_inout + size_t(b) * N, &f, &st);
      if (f_out) f_out[b] = f;
      if (stats) write_stats(st, stats + size_t(b) * 12);
    }
  };
  if (nthreads == 1) { worker(0); return; }
  std::vector<std::thread> pool;
  for (int t = 0; t < nthreads; ++t) pool.emplace_back(worker, t);
  for (std::thread& t : pool) t.join();
}

}  // namespace

#define ORACLE_TRY try {
#define ORACLE_CATCH } catch (const std::exception& e) { g_error = e.what(); return 1; } return 0;

extern "C" {

const char* oracle_last_error() { return g_error.c_str(); }

int oracle_problem_create(const char* name, int horizon, double alpha, int step_num, void** out) {
  ORACLE_TRY
  auto op = std::make_unique<OracleProblem>();
  op->name = name;
  op->ocp = ocp_problems::make_problem(name, ocp_problems::default_yaml(name, horizon, alpha, step_num, false));
  casadi::SXDict nlp = op->ocp->getNlp();
  op->ref64 = std::make_unique<SqpReference<double>>(nlp, step_num, alpha);
  op->ref32 = std::make_unique<SqpReference<float>>(nlp, step_num, alpha);
  op->nf = op->ocp->OCPConfigPtr_->getFrameSize();
  op->horizon = op->ocp->OCPConfigPtr_->getHorizon();
  op->lbx = densify(casadi::DM::vertcat(op->ocp->OCPConfigPtr_->getLowerBounds())).nonzeros();
  op->ubx = densify(casadi::DM::vertcat(op->ocp->OCPConfigPtr_->getUpperBounds())).nonzeros();
  op->lbg = densify(casadi::DM::vertcat(op->ocp->getConstraintLowerBounds())).nonzeros();
  op->ubg = densify(casadi::DM::vertcat(op->ocp->getConstraintUpperBounds())).nonzeros();
  *out = op.release();
  ORACLE_CATCH
}

int oracle_problem_destroy(void* h) { delete static_cast<OracleProblem*>(h); return 0; }

// dims: np nf horizon ng n m nnz_h nnz_a
int oracle_problem_dims(void* h, int* dims) {
  OracleProblem* op = static_cast<OracleProblem*>(h);
  const auto& r = *op->ref64;
  dims[0] = r.np(); dims[1] = op->nf; dims[2] = op->horizon; dims[3] = r.ng(); dims[4] = r.n(); dims[5] = r.m();
  dims[6] = static_cast<int>(r.h_rowidx().size()); dims[7] = static_cast<int>(r.a_rowidx().size());
  return 0;
}

int oracle_problem_patterns(void* h, int* hp, int* hi, int* ap, int* ai) {
  const auto& r = *static_cast<OracleProblem*>(h)->ref64;
  std::memcpy(hp, r.h_colptr().data(), r.h_colptr().size() * sizeof(int));
  std::memcpy(hi, r.h_rowidx().data(), r.h_rowidx().size() * sizeof(int));
  std::memcpy(ap, r.a_colptr().data(), r.a_colptr().size() * sizeof(int));
  std::memcpy(ai, r.a_rowidx().data(), r.a_rowidx().size() * sizeof(int));
  return 0;
}

int oracle_problem_bounds(void* h, double* lbx, double* ubx, double* lbg, double* ubg) {
  OracleProblem* op = static_cast<OracleProblem*>(h);
  std::memcpy(lbx, op->lbx.data(), op->lbx.size() * 8); std::memcpy(ubx, op->ubx.data(), op->ubx.size() * 8);
  std::memcpy(lbg, op->lbg.data(), op->lbg.size() * 8); std::memcpy(ubg, op->ubg.data(), op->ubg.size() * 8);
  return 0;
}

int oracle_problem_set_qp_settings(void* h, const double* settings) {
  OracleProblem* op = static_cast<OracleProblem*>(h);
  apply_settings(settings, op->ref64->qp_settings);
  apply_settings(settings, op->ref32->qp_settings);
  return 0;
}

int oracle_problem_set_schedule(void* h, int step_num, double alpha, int reuse_symbolic) {
  OracleProblem* op = static_cast<OracleProblem*>(h);
  op->ref64->setSchedule(step_num, alpha); op->ref32->setSchedule(step_num, alpha);
  op->ref64->qp_settings.reuse_symbolic = reuse_symbolic != 0;
  op->ref32->qp_settings.reuse_symbolic = reuse_symbolic != 0;
  return 0;
}

// local system of ONE instance at x, first frame pinned when frame != NULL
int oracle_local_system(void* h, const double* frame, const double* p, const double* x, double* hv, double* q,
                        double* av, double* l, double* u) {
  ORACLE_TRY
  OracleProblem* op = static_cast<OracleProblem*>(h);
  const auto& r = *op->ref64;
  SqpReference<double>::Work k;
  r.init_work(k);
  std::vector<double> lbx(op->lbx), ubx(op->ubx);
  if (frame) for (int i = 0; i < op->nf; ++i) { lbx[i] = frame[i]; ubx[i] = frame[i]; }
  r.local_system(k, p, x, lbx.data(), ubx.data(), op->lbg.data(), op->ubg.data());
  std::memcpy(hv, k.hv.data(), k.hv.size() * 8); std::memcpy(q, k.q.data(), k.q.size() * 8);
  std::memcpy(av, k.av.data(), k.av.size() * 8); std::memcpy(l, k.l.data(), k.l.size() * 8);
  std::memcpy(u, k.u.data(), k.u.size() * 8);
  ORACLE_CATCH
}

int oracle_objective(void* h, const double* p, const double* x, double* f) {
  ORACLE_TRY
  OracleProblem* op = static_cast<OracleProblem*>(h);
  SqpReference<double>::Work k;
  *f = op->ref64->objective(k, p, x);
  ORACLE_CATCH
}

// restated reference CPU path for B instances, `nthreads` host threads, FP64 or FP32 QP
int oracle_sqp_solve_batch(void* h, int B, const double* frames, const double* p, double* x_inout, double* f_out,
                           double* stats, int nthreads, int use_float) {
  ORACLE_TRY
  OracleProblem* op = static_cast<OracleProblem*>(h);
  if (use_float) sqp_batch_impl<float>(*op, *op->ref32, B, frames, p, x_inout, f_out, stats, nthreads);
  else sqp_batch_impl<double>(*op, *op->ref64, B, frames, p, x_inout, f_out, stats, nthreads);
  ORACLE_CATCH
}

int oracle_qp_solve(int n, int m, const int* hp, const int* hi, const double* hx, const double* q, const int* ap,
                    const int* ai, const double* ax, const double* l, const double* u, const double* settings,
                    double* x_out, double* y_out, double* info_out, double* trace_out, int max_trace, int* n_trace,
                    int use_float) {
  ORACLE_TRY
  int rc = use_float ? qp_solve_impl<float>(n, m, hp, hi, hx, q, ap, ai, ax, l, u, settings, x_out, y_out, info_out, trace_out, max_trace, n_trace)
                     : qp_solve_impl<double>(n, m, hp, hi, hx, q, ap, ai, ax, l, u, settings, x_out, y_out, info_out, trace_out, max_trace, n_trace);
  if (rc) return rc;
  ORACLE_CATCH
}

// test/test.cpp case `id` through the restated SQP driver.  x_out has room for 8 doubles.
int oracle_kat_solve(int id, int step_num, double alpha, int use_float, double* x_out, int* n_out, double* f_out,
                     double* stats) {
  ORACLE_TRY
  ocp_problems::KatCase kc = ocp_problems::make_kat(id);
  auto dense = [](const casadi::DM& d) { return densify(d).nonzeros(); };
  std::vector<double> lbx = dense(kc.arg.at("lbx")), ubx = dense(kc.arg.at("ubx")), lbg = dense(kc.arg.at("lbg")),
                      ubg = dense(kc.arg.at("ubg")), p = dense(kc.arg.at("p"));
  SqpStats st;
  double f = 0;
  std::vector<double> x(lbx.size(), 0.0);
  if (use_float) {
    SqpReference<float> ref(kc.nlp, step_num, alpha);
    SqpReference<float>::Work k; ref.init_work(k);
    ref.solve(k, p.data(), lbx.data(), ubx.data(), lbg.data(), ubg.data(), x.data(), &f, &st);
  } else {
    SqpReference<double> ref(kc.nlp, step_num, alpha);
    SqpReference<double>::Work k; ref.init_work(k);
    ref.solve(k, p.data(), lbx.data(), ubx.data(), lbg.data(), ubg.data(), x.data(), &f, &st);
  }
  for (size_t i = 0; i < x.size(); ++i) x_out[i] = x[i];
  *n_out = static_cast<int>(x.size());
  if (f_out) *f_out = f;
  if (stats) write_stats(st, stats);
  ORACLE_CATCH
}

// CCS patterns (casadi-lite) of the local system of test/test.cpp case `id`: sizes = {n, m, nnz_h, nnz_a};
// the index arrays are written when non-null (hp / ap: n + 1 entries, hi / ai: nnz entries, at most 64 each).
int oracle_kat_patterns(int id, int* sizes, int* hp, int* hi, int* ap, int* ai) {
  ORACLE_TRY
  ocp_problems::KatCase kc = ocp_problems::make_kat(id);
  SqpReference<double> ref(kc.nlp, 1, 1.0);
  sizes[0] = ref.n(); sizes[1] = ref.m();
  sizes[2] = static_cast<int>(ref.h_rowidx().size()); sizes[3] = static_cast<int>(ref.a_rowidx().size());
  if (hp) std::copy(ref.h_colptr().begin(), ref.h_colptr().end(), hp);
  if (hi) std::copy(ref.h_rowidx().begin(), ref.h_rowidx().end(), hi);
  if (ap) std::copy(ref.a_colptr().begin(), ref.a_colptr().end(), ap);
  if (ai) std::copy(ref.a_rowidx().begin(), ref.a_rowidx().end(), ai);
  ORACLE_CATCH
}

int oracle_kat_expected(int id, double* x_out, int* n_out) {
  ORACLE_TRY
  ocp_problems::KatCase kc = ocp_problems::make_kat(id);
  for (size_t i = 0; i < kc.expected.size(); ++i) x_out[i] = kc.expected[i];
  *n_out = static_cast<int>(kc.expected.size());
  ORACLE_CATCH
}

int oracle_sample_inputs(const char* name, int B, unsigned long long seed, double* frames, double* refs) {
  ORACLE_TRY
  std::vector<double> f, r;
  ocp_problems::sample_inputs(name, B, seed, f, r);
  std::memcpy(frames, f.data(), f.size() * 8);
  std::memcpy(refs, r.data(), r.size() * 8);
  ORACLE_CATCH
}

int oracle_hardware_threads() { return static_cast<int>(std::thread::hardware_concurrency()); }

}  // extern "C"
