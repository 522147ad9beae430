// OptimalControlProblem, solver half: genSolver() and the computeOptimalTrajectory() family
// (reference src/OptimalControlProblem.cpp:224-442 CUDA_SQP branch, :78-222).  Kept in its own
// translation unit so that the front-end half (registration, bounds, nlp assembly) links
// without the device library -- the CPU oracle uses only that half.
#include "optimal_control_problem/OptimalControlProblem.h"

#include <algorithm>
#include <filesystem>

using casadi::DM;
using casadi::SX;

void OptimalControlProblem::genSolver() {
  casadi::SXDict nlp = getNlp();
  const std::string codeDir = std::filesystem::absolute(packagePath_ + "/code_gen").string();
  if (!checkDirectoryPermissions(codeDir))
    throw std::runtime_error("Cannot create or write to code generation directory: " + codeDir);
  try {
    casadi::Dict opts;
    opts["qpsol"] = "cuda_sqp";
    opts["max_iter"] = solverSettings.SQP_settings.stepNum;
    opts["alpha"] = solverSettings.SQP_settings.alpha;
    opts["verbose"] = solverSettings.verbose ? 1 : 0;
    opts["jit"] = false;
    opts["code_dir"] = codeDir;
    opts["name"] = problemName_;
    opts["nf"] = OCPConfigPtr_->getFrameSize();
    opts["horizon"] = OCPConfigPtr_->getHorizon();
    OSQPSolverPtr_ = std::make_shared<SQPOptimizationSolver>(nlp, opts);

    if (solverSettings.genCode) {
      // the reference serialises localSystemFunction here (:403-425)
      const std::string target = codeDir + "/localSystemFunction.casadi";
      OSQPSolverPtr_->getSXLocalSystemFunction().save(target);
      // ... and, in the layout of CasADi's C code generator, as a file that CasadiCInterop.cpp (or any consumer of
      // CasADi-generated C) can compile: localSystemFunction + objective
      casadi::CodeGenerator cg(codeDir + "/localSystemFunction.c");
      cg.add(OSQPSolverPtr_->getSXLocalSystemFunction());
      cg.add(OSQPSolverPtr_->getObjectiveFunction());
      cg.generate();
      if (solverSettings.verbose) std::cout << "LocalSystemFunction saved to: " << target << std::endl;
    }
    if (solverSettings.verbose) {
      std::cout << "Problem dimensions:\nVariables: " << nlp["x"].size1() << "\nConstraints: "
                << nlp["g"].size1() << "\nParameters: " << reference_.size1() << std::endl;
    }
  } catch (const std::exception& e) {
    throw std::runtime_error("Failed to generate solver: " + std::string(e.what()));
  }
}

void OptimalControlProblem::computeOptimalTrajectory(const DM& frame, const DM& reference) {
  casadi::DMDict arg = buildSolverArguments(frame, reference);
  if (!solverInputCheck(arg)) throw std::runtime_error("Solver input validation failed");
  try {
    if (!OSQPSolverPtr_) throw std::runtime_error("genSolver() has not been called");
    casadi::DMDict res = OSQPSolverPtr_->getOptimalSolution(arg);
    firstTime_ = false;
    if (res.empty()) throw std::runtime_error("Solver returned empty result");
    optimalTrajectory_ = res.at("x");
    if (solverSettings.verbose) {
      std::cout << "\n=================== result ===================" << std::endl;
      std::cout << "objective: " << res.at("f") << std::endl;
      std::cout << "solution: " << res.at("x") << std::endl;
    }
  } catch (const std::exception& e) {
    throw std::runtime_error("Optimization failed: " + std::string(e.what()));
  }
}

const std::vector<double>& OptimalControlProblem::computeOptimalTrajectoryBatch(
    int B, const std::vector<double>& frames, const std::vector<double>& references) {
  const int nf = OCPConfigPtr_->getFrameSize();
  const int N = nf * OCPConfigPtr_->getHorizon();
  const int np = static_cast<int>(reference_.size1());
  if (B <= 0) throw std::invalid_argument("batch size must be positive");
  if (static_cast<long long>(frames.size()) != static_cast<long long>(B) * nf)
    throw std::invalid_argument("State dimension mismatch in batch");
  if (static_cast<long long>(references.size()) != static_cast<long long>(B) * np)
    throw std::invalid_argument("Reference dimension mismatch in batch");
  try {
    if (!OSQPSolverPtr_) throw std::runtime_error("genSolver() has not been called");
    if (batchSize_ != B) {  // first call (or new batch size): every instance starts from x = 0
      batchTrajectory_.assign(static_cast<size_t>(B) * N, 0.0);
      batchSize_ = B;
    }
    DM lbx = DM::vertcat(OCPConfigPtr_->getLowerBounds());
    DM ubx = DM::vertcat(OCPConfigPtr_->getUpperBounds());
    DM lbg = DM::vertcat(getConstraintLowerBounds());
    DM ubg = DM::vertcat(getConstraintUpperBounds());
    OSQPSolverPtr_->getOptimalSolutionBatch(B, frames, references, lbx, ubx, lbg, ubg, batchTrajectory_,
                                            batchObjective_, &batchStats_);
  } catch (const std::exception& e) {
    throw std::runtime_error("Optimization failed: " + std::string(e.what()));
  }
  return batchTrajectory_;
}

// Receding-horizon shift of the stored batch iterates (MPC tick loop): frame k <- frame k+1 for
// k < horizon-1, the last frame is repeated.  The reference keeps result_["x"] as the last solve
// left it (SQPOptimizationSolver.cpp:215); this is the opt-in extension for a moving horizon.
void OptimalControlProblem::shiftBatchTrajectory() {
  const size_t nf = OCPConfigPtr_->getFrameSize();
  const size_t N = nf * OCPConfigPtr_->getHorizon();
  if (N <= nf) return;
  for (int i = 0; i < batchSize_; ++i) {
    double* xi = batchTrajectory_.data() + static_cast<size_t>(i) * N;
    std::copy(xi + nf, xi + N, xi);
  }
}

void OptimalControlProblem::resetWarmStart() {
  firstTime_ = true;
  batchSize_ = 0;
  batchTrajectory_.clear();
  if (OSQPSolverPtr_) OSQPSolverPtr_->resetIterate();
}

