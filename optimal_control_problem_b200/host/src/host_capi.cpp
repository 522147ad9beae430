// host_capi.cpp -- C entry points over the C++ front-end, for ctypes (tests, bench.py,
// the Python mirror in optimal_control_problem_b200/__init__.py).  Everything here goes
// through the same classes a C++ caller uses (OptimalControlProblem, SQPOptimizationSolver,
// CuCaQP); the numerical work happens behind include/ocp_b200.h on the GPU.
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <string>

#include "optimal_control_problem/OptimalControlProblem.h"
#include "optimal_control_problem/sqp_solver/CuCaQP.h"
#include "optimal_control_problem/codegen/StageCodegen.h"
#include "problems/problems.h"

namespace {
thread_local std::string g_err;

struct HostProblem {
  std::string name;
  std::unique_ptr<OptimalControlProblem> ocp;
  std::vector<double> lbx, ubx, lbg, ubg;
};

struct HostNlp {  // a bare SQPOptimizationSolver over a test/test.cpp NLP
  ocp_problems::KatCase kat;
  std::unique_ptr<SQPOptimizationSolver> solver;
};

std::vector<double> dense(const casadi::DM& d) { return densify(d).nonzeros(); }

void cache_bounds(HostProblem* hp) {
  hp->lbx = dense(casadi::DM::vertcat(hp->ocp->OCPConfigPtr_->getLowerBounds()));
  hp->ubx = dense(casadi::DM::vertcat(hp->ocp->OCPConfigPtr_->getUpperBounds()));
  hp->lbg = dense(casadi::DM::vertcat(hp->ocp->getConstraintLowerBounds()));
  hp->ubg = dense(casadi::DM::vertcat(hp->ocp->getConstraintUpperBounds()));
}

// An OptimalControlProblem whose costs and constraints are registered from outside (the Python
// front-end of optimal_control_problem_b200/symbolic.py) instead of in an overridden
// deployConstraintsAndAddCost(): what the reference's pybind trampoline class would have done
// (src/pybind/python_bindings.cpp:380-445, commented out upstream).
class ScriptedOCP : public OptimalControlProblem {
 public:
  ScriptedOCP(YAML::Node node, const std::string& name) : OptimalControlProblem(node) { setProblemName(name); }
  void deployConstraintsAndAddCost() override {}
};

int op_code(const std::string& op, bool binary) {
  static const std::map<std::string, int> bin = {{"add", casadi::OP_ADD}, {"sub", casadi::OP_SUB}, {"mul", casadi::OP_MUL},
      {"div", casadi::OP_DIV}, {"pow", casadi::OP_POW}, {"atan2", casadi::OP_ATAN2}, {"fmin", casadi::OP_FMIN},
      {"fmax", casadi::OP_FMAX}};
  static const std::map<std::string, int> un = {{"neg", casadi::OP_NEG}, {"sq", casadi::OP_SQ}, {"sqrt", casadi::OP_SQRT},
      {"sin", casadi::OP_SIN}, {"cos", casadi::OP_COS}, {"tan", casadi::OP_TAN}, {"asin", casadi::OP_ASIN},
      {"acos", casadi::OP_ACOS}, {"atan", casadi::OP_ATAN}, {"exp", casadi::OP_EXP}, {"log", casadi::OP_LOG},
      {"fabs", casadi::OP_FABS}, {"sign", casadi::OP_SIGN}, {"tanh", casadi::OP_TANH}, {"sinh", casadi::OP_SINH},
      {"cosh", casadi::OP_COSH}};
  const auto& m = binary ? bin : un;
  auto it = m.find(op);
  if (it == m.end()) throw std::invalid_argument("unknown SX operation: " + op);
  return it->second;
}
casadi::SX& sx(void* h) {
  if (!h) throw std::invalid_argument("SX handle is NULL");
  return *static_cast<casadi::SX*>(h);
}
}  // namespace

#define HOST_TRY try {
#define HOST_CATCH } catch (const std::exception& e) { g_err = e.what(); return 1; } return 0;

extern "C" {

const char* ocp_host_last_error() { return g_err.c_str(); }

// where genSolver() caches generated stage libraries ($OCP_B200_SHARE_DIR/code_gen)
int ocp_host_set_share_dir(const char* path) { return setenv("OCP_B200_SHARE_DIR", path, 1); }

// Named benchmark problem: constructor + setReference + deployConstraintsAndAddCost + genSolver.
// genSolver() generates and nvcc-compiles the stage library; it needs no GPU.
int ocp_host_problem_create(const char* name, int horizon, double alpha, int step_num, void** out) {
  HOST_TRY
  auto hp = std::make_unique<HostProblem>();
  hp->name = name;
  hp->ocp = ocp_problems::make_problem(name, ocp_problems::default_yaml(name, horizon, alpha, step_num, false));
  hp->ocp->genSolver();
  hp->lbx = dense(casadi::DM::vertcat(hp->ocp->OCPConfigPtr_->getLowerBounds()));
  hp->ubx = dense(casadi::DM::vertcat(hp->ocp->OCPConfigPtr_->getUpperBounds()));
  hp->lbg = dense(casadi::DM::vertcat(hp->ocp->getConstraintLowerBounds()));
  hp->ubg = dense(casadi::DM::vertcat(hp->ocp->getConstraintUpperBounds()));
  *out = hp.release();
  HOST_CATCH
}

// Same, from YAML text (the `optimal_control_problem:` node or a document containing it)
int ocp_host_problem_create_yaml(const char* name, const char* yaml_text, void** out) {
  HOST_TRY
  auto hp = std::make_unique<HostProblem>();
  hp->name = name;
  hp->ocp = ocp_problems::make_problem(name, yaml_text);
  hp->ocp->genSolver();
  hp->lbx = dense(casadi::DM::vertcat(hp->ocp->OCPConfigPtr_->getLowerBounds()));
  hp->ubx = dense(casadi::DM::vertcat(hp->ocp->OCPConfigPtr_->getUpperBounds()));
  hp->lbg = dense(casadi::DM::vertcat(hp->ocp->getConstraintLowerBounds()));
  hp->ubg = dense(casadi::DM::vertcat(hp->ocp->getConstraintUpperBounds()));
  *out = hp.release();
  HOST_CATCH
}

// ---- symbolic expressions for the Python front-end: opaque casadi::SX handles ------------------------
int ocp_host_sx_sym(const char* name, int n, void** out) {
  HOST_TRY
  *out = new casadi::SX(casadi::SX::sym(name, n));
  HOST_CATCH
}
int ocp_host_sx_const(const double* v, int n, void** out) {
  HOST_TRY
  *out = new casadi::SX(std::vector<double>(v, v + n));
  HOST_CATCH
}
int ocp_host_sx_unary(const char* op, void* a, void** out) {
  HOST_TRY
  *out = new casadi::SX(casadi::SX::unary(op_code(op, false), sx(a)));
  HOST_CATCH
}
int ocp_host_sx_binary(const char* op, void* a, void* b, void** out) {
  HOST_TRY
  *out = new casadi::SX(casadi::SX::binary(op_code(op, true), sx(a), sx(b)));
  HOST_CATCH
}
int ocp_host_sx_vertcat(void** parts, int count, void** out) {
  HOST_TRY
  std::vector<casadi::SX> v;
  for (int i = 0; i < count; ++i) v.push_back(sx(parts[i]));
  *out = new casadi::SX(casadi::SX::vertcat(v));
  HOST_CATCH
}
int ocp_host_sx_slice(void* a, int start, int stop, void** out) {
  HOST_TRY
  const casadi::SX& x = sx(a);
  if (start < 0 || stop > x.size1() || start > stop) throw std::out_of_range("SX slice out of range");
  *out = new casadi::SX(x(casadi::Slice(start, stop)));
  HOST_CATCH
}
int ocp_host_sx_size(void* a) { return a ? static_cast<int>(static_cast<casadi::SX*>(a)->size1()) : -1; }
int ocp_host_sx_free(void* a) { delete static_cast<casadi::SX*>(a); return 0; }

// ---- an OptimalControlProblem assembled call by call (see ScriptedOCP) --------------------------------
int ocp_host_scripted_create(const char* name, const char* yaml_text, void** out) {
  HOST_TRY
  YAML::Node node = YAML::Load(yaml_text);
  if (node["optimal_control_problem"]) node = node["optimal_control_problem"];
  auto hp = std::make_unique<HostProblem>();
  hp->name = name;
  hp->ocp = std::make_unique<ScriptedOCP>(node, name);
  *out = hp.release();
  HOST_CATCH
}
int ocp_host_scripted_info(void* h, int* horizon, double* dt, int* frame_size) {
  HOST_TRY
  auto& cfg = *static_cast<HostProblem*>(h)->ocp->OCPConfigPtr_;
  *horizon = cfg.getHorizon(); *dt = cfg.getDt(); *frame_size = cfg.getFrameSize();
  HOST_CATCH
}
int ocp_host_scripted_variable(void* h, int k, const char* field, void** out) {
  HOST_TRY
  *out = new casadi::SX(static_cast<HostProblem*>(h)->ocp->OCPConfigPtr_->getVariable(k, field));
  HOST_CATCH
}
int ocp_host_scripted_set_reference(void* h, void* ref) {
  HOST_TRY
  static_cast<HostProblem*>(h)->ocp->setReference(sx(ref));
  HOST_CATCH
}
int ocp_host_scripted_add_scalar_cost(void* h, void* cost) {
  HOST_TRY
  static_cast<HostProblem*>(h)->ocp->addScalarCost(sx(cost));
  HOST_CATCH
}
int ocp_host_scripted_add_vector_cost(void* h, const double* w, int n, void* cost) {
  HOST_TRY
  static_cast<HostProblem*>(h)->ocp->addVectorCost(std::vector<double>(w, w + n), sx(cost));
  HOST_CATCH
}
int ocp_host_scripted_add_inequality(void* h, const char* name, const double* lb, void* expr, const double* ub, int n) {
  HOST_TRY
  static_cast<HostProblem*>(h)->ocp->addInequalityConstraint(name, casadi::DM(std::vector<double>(lb, lb + n)), sx(expr),
                                                             casadi::DM(std::vector<double>(ub, ub + n)));
  HOST_CATCH
}
int ocp_host_scripted_add_equation(void* h, const char* name, void* lhs, void* rhs) {
  HOST_TRY
  if (rhs) static_cast<HostProblem*>(h)->ocp->addEquationConstraint(name, sx(lhs), sx(rhs));
  else static_cast<HostProblem*>(h)->ocp->addEquationConstraint(name, sx(lhs));
  HOST_CATCH
}
// genSolver(): symbolic AD, stage code generation, nvcc -- no GPU needed
int ocp_host_scripted_gen_solver(void* h) {
  HOST_TRY
  HostProblem* hp = static_cast<HostProblem*>(h);
  hp->ocp->genSolver();
  cache_bounds(hp);
  HOST_CATCH
}
// YAML text of the built-in benchmark problems (so that a scripted problem can share their settings)
int ocp_host_default_yaml(const char* name, int horizon, double alpha, int step_num, char* buf, int cap) {
  HOST_TRY
  const std::string y = ocp_problems::default_yaml(name, horizon, alpha, step_num, false);
  if (static_cast<int>(y.size()) + 1 > cap) throw std::length_error("default_yaml: buffer too small");
  std::memcpy(buf, y.c_str(), y.size() + 1);
  HOST_CATCH
}

int ocp_host_problem_destroy(void* h) { delete static_cast<HostProblem*>(h); return 0; }

// dims: np nf horizon ng n m nnz_h nnz_a
int ocp_host_problem_dims(void* h, int* dims) {
  HOST_TRY
  HostProblem* hp = static_cast<HostProblem*>(h);
  auto s = hp->ocp->getSolver();
  casadi::Function f = s->getSXLocalSystemFunction();
  dims[0] = s->numParameters(); dims[1] = hp->ocp->OCPConfigPtr_->getFrameSize();
  dims[2] = hp->ocp->OCPConfigPtr_->getHorizon(); dims[3] = s->numConstraints() - s->numVariables();
  dims[4] = s->numVariables(); dims[5] = s->numConstraints();
  dims[6] = static_cast<int>(f.nnz_out(0)); dims[7] = static_cast<int>(f.nnz_out(2));
  HOST_CATCH
}

int ocp_host_problem_patterns(void* h, int* hp_, int* hi, int* ap, int* ai) {
  HOST_TRY
  casadi::Function f = static_cast<HostProblem*>(h)->ocp->getSolver()->getSXLocalSystemFunction();
  const casadi::Sparsity& hs = f.sparsity_out(0);
  const casadi::Sparsity& as = f.sparsity_out(2);
  for (size_t k = 0; k < hs.get_colind().size(); ++k) hp_[k] = static_cast<int>(hs.get_colind()[k]);
  for (size_t k = 0; k < hs.get_row().size(); ++k) hi[k] = static_cast<int>(hs.get_row()[k]);
  for (size_t k = 0; k < as.get_colind().size(); ++k) ap[k] = static_cast<int>(as.get_colind()[k]);
  for (size_t k = 0; k < as.get_row().size(); ++k) ai[k] = static_cast<int>(as.get_row()[k]);
  HOST_CATCH
}

int ocp_host_problem_bounds(void* h, double* lbx, double* ubx, double* lbg, double* ubg) {
  HostProblem* hp = static_cast<HostProblem*>(h);
  std::memcpy(lbx, hp->lbx.data(), hp->lbx.size() * 8); std::memcpy(ubx, hp->ubx.data(), hp->ubx.size() * 8);
  std::memcpy(lbg, hp->lbg.data(), hp->lbg.size() * 8); std::memcpy(ubg, hp->ubg.data(), hp->ubg.size() * 8);
  return 0;
}

const char* ocp_host_problem_model_library(void* h) {
  return static_cast<HostProblem*>(h)->ocp->getSolver()->modelLibrary().c_str();
}

// the ocp_b200_solver handle behind the problem (creates it: needs a GPU)
int ocp_host_problem_handle(void* h, ocp_b200_solver** out) {
  HOST_TRY
  *out = static_cast<HostProblem*>(h)->ocp->getSolver()->handle();
  HOST_CATCH
}

int ocp_host_problem_get_settings(void* h, ocp_b200_settings* out) {
  HOST_TRY
  *out = static_cast<HostProblem*>(h)->ocp->getSolver()->settings();
  HOST_CATCH
}

// QP / PCG settings only; alpha and step_num stay with the SQP driver (set_schedule)
int ocp_host_problem_set_settings(void* h, const ocp_b200_settings* in) {
  HOST_TRY
  auto s = static_cast<HostProblem*>(h)->ocp->getSolver();
  const double alpha = s->settings().sqp_alpha;
  const int steps = s->settings().sqp_step_num;
  s->settings() = *in;
  s->settings().sqp_alpha = alpha;
  s->settings().sqp_step_num = steps;
  HOST_CATCH
}

int ocp_host_problem_set_schedule(void* h, int step_num, double alpha) {
  HOST_TRY
  static_cast<HostProblem*>(h)->ocp->getSolver()->setSchedule(step_num, alpha);
  HOST_CATCH
}

// OptimalControlProblem::computeOptimalTrajectory (reference OptimalControlProblem.cpp:78-222):
// one instance, warm-started from the previous call.  x_out [N], f_out [1]
int ocp_host_compute_optimal_trajectory(void* h, const double* frame, const double* reference, double* x_out,
                                        double* f_out) {
  HOST_TRY
  HostProblem* hp = static_cast<HostProblem*>(h);
  const int nf = hp->ocp->OCPConfigPtr_->getFrameSize();
  const int np = static_cast<int>(hp->ocp->getReference().size1());
  hp->ocp->computeOptimalTrajectory(casadi::DM(std::vector<double>(frame, frame + nf)),
                                    casadi::DM(std::vector<double>(reference, reference + np)));
  std::vector<double> x = dense(hp->ocp->getOptimalTrajectory());
  std::memcpy(x_out, x.data(), x.size() * 8);
  if (f_out) *f_out = hp->ocp->getSolver()->lastObjective();
  HOST_CATCH
}

// batched sibling; trajectories persist inside the problem as the warm start of the next call
int ocp_host_compute_optimal_trajectory_batch(void* h, int B, const double* frames, const double* references,
                                              double* x_out, double* f_out, double* stats_out) {
  HOST_TRY
  HostProblem* hp = static_cast<HostProblem*>(h);
  const int nf = hp->ocp->OCPConfigPtr_->getFrameSize();
  const int np = static_cast<int>(hp->ocp->getReference().size1());
  const std::vector<double>& x = hp->ocp->computeOptimalTrajectoryBatch(
      B, std::vector<double>(frames, frames + size_t(B) * nf),
      std::vector<double>(references, references + size_t(B) * np));
  std::memcpy(x_out, x.data(), x.size() * 8);
  if (f_out) std::memcpy(f_out, hp->ocp->getBatchObjectives().data(), size_t(B) * 8);
  if (stats_out) std::memcpy(stats_out, hp->ocp->getBatchStats().data(), size_t(B) * OCP_B200_NSTATS * 8);
  HOST_CATCH
}

int ocp_host_problem_reset(void* h) {
  HOST_TRY
  static_cast<HostProblem*>(h)->ocp->resetWarmStart();
  HOST_CATCH
}

// localSystemFunction + objective of a generated problem as ONE C file in CasADi's code-generator layout
int ocp_host_problem_generate_c(void* h, const char* path) {
  HOST_TRY
  auto solver = static_cast<HostProblem*>(h)->ocp->getSolver();
  if (!solver) throw std::runtime_error("genSolver() has not been called");
  casadi::CodeGenerator cg(path);
  cg.add(solver->getSXLocalSystemFunction());
  cg.add(solver->getObjectiveFunction());
  cg.generate();
  HOST_CATCH
}

// CasADi-format C file -> stage library (CasadiCInterop.cpp); the library path is copied into out[cap]
int ocp_host_compile_casadi_c(const char* c_file, const char* local_system_fn, const char* objective_fn, const char* name,
                              int nf, int horizon, const char* code_dir, char* out, int cap) {
  HOST_TRY
  const std::string so = ocp_codegen::compile_casadi_c(c_file, local_system_fn, objective_fn, name, nf, horizon, code_dir, false);
  if (static_cast<int>(so.size()) + 1 > cap) throw std::runtime_error("output buffer too small");
  std::memcpy(out, so.c_str(), so.size() + 1);
  HOST_CATCH
}

// devices of computeOptimalTrajectoryBatch (SQPOptimizationSolver::setDevices); n = 0 or 1: single device
int ocp_host_problem_set_devices(void* h, const int* devices, int n) {
  HOST_TRY
  auto solver = static_cast<HostProblem*>(h)->ocp->getSolver();
  if (!solver) throw std::runtime_error("genSolver() has not been called");
  solver->setDevices(std::vector<int>(devices, devices + (n > 0 ? n : 0)));
  HOST_CATCH
}

int ocp_host_problem_shift_batch(void* h) {
  HOST_TRY
  static_cast<HostProblem*>(h)->ocp->shiftBatchTrajectory();
  HOST_CATCH
}

int ocp_host_sample_inputs(const char* name, int B, unsigned long long seed, double* frames, double* refs) {
  HOST_TRY
  std::vector<double> f, r;
  ocp_problems::sample_inputs(name, B, seed, f, r);
  std::memcpy(frames, f.data(), f.size() * 8);
  std::memcpy(refs, r.data(), r.size() * 8);
  HOST_CATCH
}

// ---- test/test.cpp cases through SQPOptimizationSolver::getOptimalSolution --------------------
int ocp_host_kat_create(int id, int step_num, double alpha, void** out) {
  HOST_TRY
  auto hn = std::make_unique<HostNlp>();
  hn->kat = ocp_problems::make_kat(id);
  casadi::Dict opts;
  opts["max_iter"] = step_num;
  opts["alpha"] = alpha;
  opts["verbose"] = 0;
  opts["name"] = std::string("kat") + std::to_string(id);
  hn->solver = std::make_unique<SQPOptimizationSolver>(hn->kat.nlp, opts);
  *out = hn.release();
  HOST_CATCH
}

int ocp_host_kat_destroy(void* h) { delete static_cast<HostNlp*>(h); return 0; }

// x_out has room for 8 doubles
int ocp_host_kat_solve(void* h, double* x_out, int* n_out, double* f_out) {
  HOST_TRY
  HostNlp* hn = static_cast<HostNlp*>(h);
  casadi::DMDict res = hn->solver->getOptimalSolution(hn->kat.arg);
  std::vector<double> x = dense(res.at("x"));
  for (size_t i = 0; i < x.size(); ++i) x_out[i] = x[i];
  *n_out = static_cast<int>(x.size());
  if (f_out) *f_out = dense(res.at("f")).at(0);
  HOST_CATCH
}

int ocp_host_kat_set_settings(void* h, const ocp_b200_settings* in) {
  HOST_TRY
  HostNlp* hn = static_cast<HostNlp*>(h);
  const double alpha = hn->solver->settings().sqp_alpha;
  const int steps = hn->solver->settings().sqp_step_num;
  hn->solver->settings() = *in;
  hn->solver->settings().sqp_alpha = alpha;
  hn->solver->settings().sqp_step_num = steps;
  HOST_CATCH
}

// ---- CuCaQP life cycle on raw CCS arrays (reference CuCaQP.cpp:271-288, 183-224) -------------
int ocp_host_cucaqp_solve(int n, int m, const int* hp_, const int* hi, const double* hx, const double* q,
                          const int* ap, const int* ai, const double* ax, const double* l, const double* u,
                          double eps_abs, double eps_rel, int max_iter, double* x_out, double* y_out, double* info) {
  HOST_TRY
  using casadi::casadi_int;
  std::vector<casadi_int> hc(hp_, hp_ + n + 1), hr(hi, hi + hp_[n]), ac(ap, ap + n + 1), ar(ai, ai + ap[n]);
  casadi::DM P(casadi::Sparsity(n, n, hc, hr), std::vector<double>(hx, hx + hp_[n]));
  casadi::DM A(casadi::Sparsity(m, n, ac, ar), std::vector<double>(ax, ax + ap[n]));
  CuCaQP qp;
  if (!qp.setDimension(n, m)) throw std::runtime_error("setDimension failed");
  qp.setVerbosity(false);
  qp.setWarmStart(true);
  qp.setAbsoluteTolerance(eps_abs);
  qp.setRelativeTolerance(eps_rel);
  qp.setMaxIteration(max_iter);
  qp.setSystem({P, casadi::DM(std::vector<double>(q, q + n)), A, casadi::DM(std::vector<double>(l, l + m)),
                casadi::DM(std::vector<double>(u, u + m))});
  if (!qp.initSolver()) throw std::runtime_error(std::string("initSolver failed: ") + ocp_b200_last_error());
  if (!qp.solve()) throw std::runtime_error(std::string("solve failed: ") + ocp_b200_last_error());
  std::vector<double> x = qp.getSolution(), y = qp.getDualSolution();
  std::memcpy(x_out, x.data(), x.size() * 8);
  if (y_out) std::memcpy(y_out, y.data(), y.size() * 8);
  if (info) std::memcpy(info, qp.getInfo().data(), OCP_B200_NINFO * 8);
  HOST_CATCH
}

}  // extern "C"
