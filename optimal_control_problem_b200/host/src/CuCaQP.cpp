// CuCaQP implementation on top of the C ABI (include/ocp_b200.h).
// Lifecycle mirrors the reference (src/sqp_solver/CuCaQP.cpp): setDimension ->
// setSystem{P,q,A,l,u} -> initSolver -> solve -> getSolution.  The reference's
// initSolver() is osqp_setup (allocation, scaling, factorisation, cold start,
// :183-197); here the handle (index structures on the device) is kept while the
// sparsity is unchanged, and scaling + cold start happen inside the solve kernel,
// so every solve() is still a cold-started, freshly scaled QP like the reference's.
#include "optimal_control_problem/sqp_solver/CuCaQP.h"

#include <iomanip>
#include <iostream>

namespace {
std::vector<int> toInt(const std::vector<casadi::casadi_int>& v) {
  return std::vector<int>(v.begin(), v.end());
}
std::vector<double> denseColumn(const casadi::DM& v) {
  casadi::DM d = densify(v);
  return d.nonzeros();
}
}  // namespace

CuCaQP::CuCaQP()
    : solver_(nullptr), numOfVariables_(0), numOfConstraints_(0), isInitialized_(false), verbose_(false) {
  ocp_b200_default_settings(&settings_);
}

CuCaQP::~CuCaQP() { clearSolver(); }

void CuCaQP::clearSolver() {
  if (solver_) ocp_b200_destroy(solver_);
  solver_ = nullptr;
  isInitialized_ = false;
}

bool CuCaQP::setDimension(int numOfVariables, int numOfConstraints) {
  if (numOfVariables <= 0 || numOfConstraints <= 0) {
    std::cerr << "Error: Invalid dimensions." << std::endl;
    return false;
  }
  clearSolver();
  numOfVariables_ = numOfVariables;
  numOfConstraints_ = numOfConstraints;
  return true;
}

bool CuCaQP::setHessianMatrix(const casadi::DM& hessian) {
  if (hessian.size1() != numOfVariables_ || hessian.size2() != numOfVariables_) {
    std::cerr << "Error: Hessian matrix dimensions mismatch. Expected " << numOfVariables_ << "x"
              << numOfVariables_ << std::endl;
    return false;
  }
  if (isInitialized_ && hessian.sparsity() != hessianMatrix.sparsity()) clearSolver();
  hessianMatrix = hessian;
  return true;
}

bool CuCaQP::setLinearConstraintsMatrix(const casadi::DM& A) {
  if (A.size1() != numOfConstraints_ || A.size2() != numOfVariables_) {
    std::cerr << "Error: Constraint matrix dimensions mismatch. Expected " << numOfConstraints_ << "x"
              << numOfVariables_ << std::endl;
    return false;
  }
  if (isInitialized_ && A.sparsity() != linearConstraintMatrix.sparsity()) clearSolver();
  linearConstraintMatrix = A;
  return true;
}

bool CuCaQP::setGradient(const std::vector<OSQPFloat>& q) {
  if (static_cast<int>(q.size()) != numOfVariables_) {
    std::cerr << "Error: Gradient vector size mismatch. Expected " << numOfVariables_ << std::endl;
    return false;
  }
  gradient = q;
  return true;
}
bool CuCaQP::setLowerBound(const std::vector<OSQPFloat>& l) {
  if (static_cast<int>(l.size()) != numOfConstraints_) {
    std::cerr << "Error: Lower bound vector size mismatch. Expected " << numOfConstraints_ << std::endl;
    return false;
  }
  lowerBound = l;
  return true;
}
bool CuCaQP::setUpperBound(const std::vector<OSQPFloat>& u) {
  if (static_cast<int>(u.size()) != numOfConstraints_) {
    std::cerr << "Error: Upper bound vector size mismatch. Expected " << numOfConstraints_ << std::endl;
    return false;
  }
  upperBound = u;
  return true;
}
bool CuCaQP::setGradient(const casadi::DM& q) { return setGradient(denseColumn(q)); }
bool CuCaQP::setLowerBound(const casadi::DM& l) { return setLowerBound(denseColumn(l)); }
bool CuCaQP::setUpperBound(const casadi::DM& u) { return setUpperBound(denseColumn(u)); }

void CuCaQP::setVerbosity(bool verbosity) { verbose_ = verbosity; }
void CuCaQP::setWarmStart(bool) {
  // The reference switches OSQP warm start on (SQPOptimizationSolver.cpp:82) but destroys the
  // solver before every solve (CuCaQP.cpp:273-276), so every QP starts from x = z = y = 0.
}
void CuCaQP::setAbsoluteTolerance(OSQPFloat tolerance) { settings_.eps_abs = tolerance; }
void CuCaQP::setRelativeTolerance(OSQPFloat tolerance) { settings_.eps_rel = tolerance; }
void CuCaQP::setMaxIteration(int maxIteration) { settings_.admm_max_iter = maxIteration; }

void CuCaQP::setSystem(casadi::DMVector localSystem) {
  // same order as the reference (CuCaQP.cpp:283-287); like there, failures are only reported
  setHessianMatrix(localSystem.at(0));
  setGradient(localSystem.at(1));
  setLinearConstraintsMatrix(localSystem.at(2));
  setLowerBound(localSystem.at(3));
  setUpperBound(localSystem.at(4));
}

bool CuCaQP::initSolver() {
  if (numOfVariables_ <= 0 || hessianMatrix.size1() != numOfVariables_ ||
      linearConstraintMatrix.size1() != numOfConstraints_) {
    std::cerr << "Error: Failed to initialize solver." << std::endl;
    return false;
  }
  if (solver_) {
    if (ocp_b200_update_settings(solver_, &settings_) != OCP_B200_OK) return false;
    isInitialized_ = true;
    return true;
  }
  std::vector<int> hc = toInt(hessianMatrix.sparsity().get_colind());
  std::vector<int> hr = toInt(hessianMatrix.sparsity().get_row());
  std::vector<int> ac = toInt(linearConstraintMatrix.sparsity().get_colind());
  std::vector<int> ar = toInt(linearConstraintMatrix.sparsity().get_row());
  ocp_b200_problem_desc desc{};
  desc.np = 0;
  desc.nf = numOfVariables_;
  desc.horizon = 1;
  desc.ng = numOfConstraints_ - numOfVariables_;
  desc.nnz_h = static_cast<int>(hr.size()); desc.h_colptr = hc.data(); desc.h_rowidx = hr.data();
  desc.nnz_a = static_cast<int>(ar.size()); desc.a_colptr = ac.data(); desc.a_rowidx = ar.data();
  desc.model_library = nullptr;
  desc.num_blocks = 0; desc.block_ptr = nullptr;
  desc.device = 0;
  int rc = ocp_b200_create(&desc, &settings_, &solver_);
  if (rc != OCP_B200_OK) {
    std::cerr << "Error: Failed to initialize solver: " << ocp_b200_last_error() << std::endl;
    solver_ = nullptr;
    return false;
  }
  isInitialized_ = true;
  return true;
}

bool CuCaQP::solve() {
  if (!isInitialized_) {
    std::cerr << "Error: Solver not initialized. Call initSolver() first." << std::endl;
    return false;
  }
  solution_.assign(numOfVariables_, 0.0);
  dual_.assign(numOfConstraints_, 0.0);
  info_.assign(OCP_B200_NINFO, 0.0);
  int rc = ocp_b200_qp_solve_batch(solver_, 1, hessianMatrix.nonzeros().data(), gradient.data(),
                                   linearConstraintMatrix.nonzeros().data(), lowerBound.data(),
                                   upperBound.data(), solution_.data(), dual_.data(), info_.data());
  if (rc != OCP_B200_OK) {
    std::cerr << "Error: Failed to solve problem. Error code: " << rc << " (" << ocp_b200_last_error() << ")"
              << std::endl;
    return false;
  }
  if (verbose_)
    std::cout << "QP status " << info_[OCP_B200_INFO_STATUS] << ", ADMM iterations "
              << info_[OCP_B200_INFO_ITERS] << ", prim_res " << info_[OCP_B200_INFO_PRIM_RES]
              << ", dual_res " << info_[OCP_B200_INFO_DUAL_RES] << std::endl;
  return true;  // like the reference, the QP status itself is not inspected (CuCaQP.cpp:205-209)
}

std::vector<OSQPFloat> CuCaQP::getSolution() { return solution_; }
std::vector<OSQPFloat> CuCaQP::getDualSolution() { return dual_; }
casadi::DM CuCaQP::getSolutionAsDM() { return casadi::DM(solution_); }

void CuCaQP::printSolverData() {
  auto dump = [](const char* tag, const std::vector<OSQPFloat>& v) {
    std::cout << tag << ":\n";
    for (double x : v) std::cout << x << " ";
    std::cout << std::endl;
  };
  dump("q", gradient);
  dump("l", lowerBound);
  dump("u", upperBound);
  auto dumpMat = [](const char* tag, const casadi::DM& M) {
    std::cout << tag << " (nonzeros):\n";
    const casadi::casadi_int* ci = M.sparsity().colind();
    const casadi::casadi_int* ri = M.sparsity().row();
    for (casadi::casadi_int j = 0; j < M.size2(); ++j)
      for (casadi::casadi_int k = ci[j]; k < ci[j + 1]; ++k)
        std::cout << "(" << ri[k] << "," << j << "): " << M.nonzeros()[k] << std::endl;
  };
  dumpMat("P", hessianMatrix);
  dumpMat("A", linearConstraintMatrix);
}
