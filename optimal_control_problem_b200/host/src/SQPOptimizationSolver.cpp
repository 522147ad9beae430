// SQPOptimizationSolver implementation.
// Constructor: the symbolic set-up of the reference, step for step
// (src/sqp_solver/SQPOptimizationSolver.cpp:12-92): augmented variables w = [p; x],
// augmented constraints c = [p; x; g], H = hess_w f, grad = grad_w f, J = dc/dw,
// l' = l - c, u' = u - c.  Then, instead of keeping a CasADi virtual machine for the hot
// loop, it emits the local system as CUDA stage functions, compiles them with nvcc and
// creates the device solver.  getOptimalSolution (reference :127-216) is one call into
// the C ABI; nothing of the hot loop runs on the host.
#include "optimal_control_problem/sqp_solver/SQPOptimizationSolver.h"

#include <cmath>
#include <cstdlib>
#include <iomanip>
#include <iostream>
#include <sstream>

#include "optimal_control_problem/codegen/StageCodegen.h"

using namespace casadi;

namespace {
std::vector<int> toInt(const std::vector<casadi_int>& v) { return std::vector<int>(v.begin(), v.end()); }
std::vector<double> denseColumn(const DM& v) { return densify(v).nonzeros(); }
void check(int rc, const char* what) {
  if (rc != OCP_B200_OK)
    throw std::runtime_error(std::string(what) + " failed (" + std::to_string(rc) + "): " + ocp_b200_last_error());
}
}  // namespace

SQPOptimizationSolver::SQPOptimizationSolver(SXDict& nlp, Dict& options) : verbose_(false) {
  stepNum_ = static_cast<int>(options.at("max_iter").as_int());
  alpha_ = options.at("alpha").as_double();
  setVerbose(options.at("verbose"));

  if (nlp.find("f") == nlp.end()) throw std::invalid_argument("objective 'f' is not defined");
  if (nlp.find("x") == nlp.end()) throw std::invalid_argument("decision variables 'x' are not defined");
  SX objectExpr = nlp["f"];
  SX variables = nlp["x"];
  SX constraints = nlp.find("g") != nlp.end() ? nlp["g"] : SX();
  SX reference = nlp.find("p") != nlp.end() ? nlp["p"] : SX();

  objectiveFunction_ = Function("objective", {reference, variables}, {objectExpr});

  SX augmentedVariables = SX::vertcat({reference, variables});
  objectiveFunctionAutoDifferentiatorPtr_ = std::make_shared<AutoDifferentiator>(augmentedVariables, objectExpr);
  SX augmentedConstraints = SX::vertcat({reference, variables, constraints});
  constraintsAutoDifferentiator_ = std::make_shared<AutoDifferentiator>(augmentedVariables, augmentedConstraints);

  SX hessian = objectiveFunctionAutoDifferentiatorPtr_->getHessian(augmentedVariables);
  SX gradient = objectiveFunctionAutoDifferentiatorPtr_->getGradient(augmentedVariables);
  SXVector linearized = constraintsAutoDifferentiator_->getLinearization(augmentedVariables);
  casadi_int numOfConstraints = linearized[1].size1();

  SX l = SX::sym("l", numOfConstraints);
  SX u = SX::sym("u", numOfConstraints);
  SX l_linearized = l + linearized[1];
  SX u_linearized = u + linearized[1];

  localSystemFunction_ = Function("localSystemFunction", {reference, variables, l, u},
                                  {hessian, gradient, linearized[0], l_linearized, u_linearized});

  np_ = static_cast<int>(reference.numel());
  N_ = static_cast<int>(variables.numel());
  ng_ = static_cast<int>(constraints.numel());
  n_ = np_ + N_;
  m_ = n_ + ng_;

  // ---- device side: stage functions + solver handle -------------------------------
  ocp_codegen::ModelSpec spec;
  spec.name = options.count("name") ? options.at("name").as_string() : std::string("nlp");
  spec.p = reference; spec.x = variables; spec.f = objectExpr; spec.g = constraints;
  spec.grad = gradient; spec.hess = hessian; spec.jac = linearized[0];
  spec.nf = options.count("nf") ? static_cast<int>(options.at("nf").as_int()) : N_;
  spec.horizon = options.count("horizon") ? static_cast<int>(options.at("horizon").as_int()) : 1;
  if (spec.nf * spec.horizon != N_) { spec.nf = N_; spec.horizon = 1; }
  std::string codeDir;
  if (options.count("code_dir")) codeDir = options.at("code_dir").as_string();
  else if (const char* e = std::getenv("OCP_B200_CODE_DIR")) codeDir = e;
  else codeDir = "./ocp_b200_share/code_gen";
  ocp_codegen::ModelSource src = ocp_codegen::generate(spec);
  modelLibrary_ = ocp_codegen::compile(src, spec.name, codeDir, verbose_);

  ocp_b200_default_settings(&settings_);
  settings_.sqp_alpha = alpha_;
  settings_.sqp_step_num = stepNum_;
  settings_.eps_abs = 1e-3;        // reference :83
  settings_.eps_rel = 1e-3;        // reference :84
  settings_.admm_max_iter = 10000; // reference :85
  if (options.count("pcg_tol")) settings_.pcg_tol = options.at("pcg_tol").as_double();
  if (options.count("pcg_max_iter")) settings_.pcg_max_iter = static_cast<int>(options.at("pcg_max_iter").as_int());
  if (options.count("pcg_precond")) settings_.pcg_precond = static_cast<int>(options.at("pcg_precond").as_int());

  hColptr_ = toInt(hessian.sparsity().get_colind()); hRowidx_ = toInt(hessian.sparsity().get_row());
  aColptr_ = toInt(linearized[0].sparsity().get_colind()); aRowidx_ = toInt(linearized[0].sparsity().get_row());
  nf_ = spec.nf; horizon_ = spec.horizon;
  device_ = options.count("device") ? static_cast<int>(options.at("device").as_int()) : 0;
  if (const char* e = std::getenv("OCP_B200_DEVICES")) {   // "0,1,2,3": devices of the batched entry point
    std::vector<int> list;
    std::stringstream ss(e);
    for (std::string tok; std::getline(ss, tok, ',');)
      if (!tok.empty()) list.push_back(std::atoi(tok.c_str()));
    devices_ = list;
  }
  // The device handle is created on first use (ensureDevice) so that the symbolic set-up and
  // the nvcc build can run on a machine without a GPU; every solve entry point needs the GPU.

  result_ = {{"x", DM::zeros(variables.size1())}, {"f", DM::zeros(1)}};
}

SQPOptimizationSolver::~SQPOptimizationSolver() {
  if (multi_) ocp_b200_destroy_multi(multi_);
  if (handle_) ocp_b200_destroy(handle_);
}

void SQPOptimizationSolver::setDevices(const std::vector<int>& devices) {
  if (devices == devices_) return;
  if (multi_) { ocp_b200_destroy_multi(multi_); multi_ = nullptr; }
  devices_ = devices;
}

void SQPOptimizationSolver::ensureDevice() {
  if (handle_) return;
  ocp_b200_problem_desc desc{};
  desc.np = np_; desc.nf = nf_; desc.horizon = horizon_; desc.ng = ng_;
  desc.nnz_h = static_cast<int>(hRowidx_.size()); desc.h_colptr = hColptr_.data(); desc.h_rowidx = hRowidx_.data();
  desc.nnz_a = static_cast<int>(aRowidx_.size()); desc.a_colptr = aColptr_.data(); desc.a_rowidx = aRowidx_.data();
  desc.model_library = modelLibrary_.c_str();
  desc.device = device_;
  check(ocp_b200_create(&desc, &settings_, &handle_), "ocp_b200_create");
}

void SQPOptimizationSolver::applySettings() {
  ensureDevice();
  check(ocp_b200_update_settings(handle_, &settings_), "ocp_b200_update_settings");
}

void SQPOptimizationSolver::setSchedule(int stepNum, double alpha) {
  stepNum_ = stepNum;
  alpha_ = alpha;
  settings_.sqp_step_num = stepNum_;
  settings_.sqp_alpha = alpha_;
  if (handle_) check(ocp_b200_update_settings(handle_, &settings_), "ocp_b200_update_settings");
}

void SQPOptimizationSolver::resetIterate() { result_["x"] = DM::zeros(N_); result_["f"] = DM::zeros(1); }

DMDict SQPOptimizationSolver::getOptimalSolution(const DMDict& arg) {
  std::vector<double> lbx = denseColumn(arg.at("lbx")), ubx = denseColumn(arg.at("ubx"));
  std::vector<double> lbg = denseColumn(arg.at("lbg")), ubg = denseColumn(arg.at("ubg"));
  std::vector<double> p;
  if (arg.find("p") != arg.end()) p = denseColumn(arg.at("p"));
  if (static_cast<int>(lbx.size()) != N_ || static_cast<int>(ubx.size()) != N_ ||
      static_cast<int>(lbg.size()) != ng_ || static_cast<int>(ubg.size()) != ng_ ||
      static_cast<int>(p.size()) != np_)
    throw std::invalid_argument("getOptimalSolution: argument dimensions do not match the nlp");
  std::vector<double> x = denseColumn(result_.at("x"));
  double f = 0.0;
  std::vector<double> stats(OCP_B200_NSTATS, 0.0);

  if (!verbose_) {
    settings_.sqp_step_num = stepNum_;
    settings_.sqp_alpha = alpha_;
    applySettings();
    check(ocp_b200_solve_batch(handle_, 1, nullptr, p.data(), lbx.data(), ubx.data(), lbg.data(),
                               ubg.data(), x.data(), &f, stats.data()), "ocp_b200_solve_batch");
  } else {
    // verbose mode steps one SQP iteration at a time so that the reference's verbose-only
    // early exit (||dx||_2 < 1e-6, SQPOptimizationSolver.cpp:183-197) is reproduced
    std::cout << "=== SQP start: steps " << stepNum_ << ", alpha " << alpha_ << " ===" << std::endl;
    settings_.sqp_step_num = 1;
    settings_.sqp_alpha = alpha_;
    applySettings();
    for (int i = 0; i < stepNum_; ++i) {
      std::vector<double> old = x;
      check(ocp_b200_solve_batch(handle_, 1, nullptr, p.data(), lbx.data(), ubx.data(), lbg.data(),
                                 ubg.data(), x.data(), &f, stats.data()), "ocp_b200_solve_batch");
      double nd = 0.0;
      for (size_t k = 0; k < x.size(); ++k) nd += (x[k] - old[k]) * (x[k] - old[k]);
      nd = std::sqrt(nd);
      std::cout << "  step " << i + 1 << "/" << stepNum_ << ": f = " << f << ", ||dx|| = " << nd
                << ", ADMM iterations " << stats[OCP_B200_STAT_LAST_ADMM] << ", QP status "
                << stats[OCP_B200_STAT_QP_STATUS] << std::endl;
      if (nd < 1e-6) { std::cout << "  converged, stopping early" << std::endl; break; }
    }
    settings_.sqp_step_num = stepNum_;
    applySettings();
  }
  result_.at("x") = DM(x);
  result_.at("f") = DM(f);
  return result_;
}

void SQPOptimizationSolver::getOptimalSolutionBatch(int B, const std::vector<double>& frames,
                                                    const std::vector<double>& p, const DM& lbx,
                                                    const DM& ubx, const DM& lbg, const DM& ubg,
                                                    std::vector<double>& x_inout, std::vector<double>& f_out,
                                                    std::vector<double>* stats) {
  std::vector<double> lx = denseColumn(lbx), ux = denseColumn(ubx), lg = denseColumn(lbg), ug = denseColumn(ubg);
  if (static_cast<int>(lx.size()) != N_ || static_cast<int>(ux.size()) != N_ ||
      static_cast<int>(lg.size()) != ng_ || static_cast<int>(ug.size()) != ng_)
    throw std::invalid_argument("getOptimalSolutionBatch: bound dimensions do not match the nlp");
  if (static_cast<long long>(p.size()) != static_cast<long long>(B) * np_ ||
      static_cast<long long>(x_inout.size()) != static_cast<long long>(B) * N_)
    throw std::invalid_argument("getOptimalSolutionBatch: batch dimensions do not match");
  f_out.assign(B, 0.0);
  if (stats) stats->assign(static_cast<size_t>(B) * OCP_B200_NSTATS, 0.0);
  settings_.sqp_step_num = stepNum_;
  settings_.sqp_alpha = alpha_;
  if (devices_.size() > 1) {   // several GPUs: contiguous blocks of the batch, one host thread per device
    if (!multi_) {
      ocp_b200_problem_desc desc{};
      desc.np = np_; desc.nf = nf_; desc.horizon = horizon_; desc.ng = ng_;
      desc.nnz_h = static_cast<int>(hRowidx_.size()); desc.h_colptr = hColptr_.data(); desc.h_rowidx = hRowidx_.data();
      desc.nnz_a = static_cast<int>(aRowidx_.size()); desc.a_colptr = aColptr_.data(); desc.a_rowidx = aRowidx_.data();
      desc.model_library = modelLibrary_.c_str();
      check(ocp_b200_create_multi(&desc, &settings_, devices_.data(), static_cast<int>(devices_.size()), &multi_),
            "ocp_b200_create_multi");
    }
    check(ocp_b200_multi_update_settings(multi_, &settings_), "ocp_b200_multi_update_settings");
    check(ocp_b200_solve_batch_multi(multi_, B, frames.empty() ? nullptr : frames.data(), p.data(), lx.data(), ux.data(),
                                     lg.data(), ug.data(), x_inout.data(), f_out.data(), stats ? stats->data() : nullptr),
          "ocp_b200_solve_batch_multi");
    return;
  }
  applySettings();
  check(ocp_b200_solve_batch(handle_, B, frames.empty() ? nullptr : frames.data(), p.data(), lx.data(),
                             ux.data(), lg.data(), ug.data(), x_inout.data(), f_out.data(),
                             stats ? stats->data() : nullptr), "ocp_b200_solve_batch");
}

DMVector SQPOptimizationSolver::getLocalSystemGPU(const DMDict& arg) {
  std::vector<double> lbx = denseColumn(arg.at("lbx")), ubx = denseColumn(arg.at("ubx"));
  std::vector<double> lbg = denseColumn(arg.at("lbg")), ubg = denseColumn(arg.at("ubg"));
  std::vector<double> p;
  if (arg.find("p") != arg.end()) p = denseColumn(arg.at("p"));
  std::vector<double> x = denseColumn(result_.at("x"));
  const Sparsity& hs = localSystemFunction_.sparsity_out(0);
  const Sparsity& as = localSystemFunction_.sparsity_out(2);
  std::vector<double> hv(hs.nnz()), q(n_), av(as.nnz()), l(m_), u(m_);
  ensureDevice();
  check(ocp_b200_export_qp(handle_, 1, nullptr, p.data(), lbx.data(), ubx.data(), lbg.data(), ubg.data(),
                           x.data(), hv.data(), q.data(), av.data(), l.data(), u.data()), "ocp_b200_export_qp");
  return {DM(hs, hv), DM(q), DM(as, av), DM(l), DM(u)};
}

Function SQPOptimizationSolver::getSXLocalSystemFunction() const { return localSystemFunction_; }

void SQPOptimizationSolver::setVerbose(bool verbose) { verbose_ = verbose; }
