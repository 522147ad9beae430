// CuCaQP: QP adapter  min 1/2 x'Px + q'x  s.t.  l <= Ax <= u.
// Public surface of the reference class (include/optimal_control_problem/sqp_solver/
// CuCaQP.h:27-102).  The reference converts CasADi DMs to Eigen<float> and drives
// OsqpEigen -> OSQP (CuCaQP.cpp:183-224); this one keeps the CCS arrays in FP64 and
// drives the sm_100a ADMM kernels through the C ABI (include/ocp_b200.h,
// ocp_b200_qp_solve_batch).  Eigen is not in this image: the Eigen overloads of
// the reference are offered on std::vector<double> instead.
#pragma once

#include <vector>

#include "casadi/casadi.hpp"
#include "ocp_b200.h"

typedef double OSQPFloat;  // the reference builds OSQP with OSQP_USE_FLOAT=ON; this path is FP64

class CuCaQP {
 public:
  CuCaQP();
  ~CuCaQP();
  CuCaQP(const CuCaQP&) = delete;
  CuCaQP& operator=(const CuCaQP&) = delete;

  bool setDimension(int numOfVariables, int numOfConstraints);

  bool setHessianMatrix(const casadi::DM& hessian);
  bool setGradient(const casadi::DM& q);
  bool setLinearConstraintsMatrix(const casadi::DM& A);
  bool setLowerBound(const casadi::DM& l);
  bool setUpperBound(const casadi::DM& u);
  bool setGradient(const std::vector<OSQPFloat>& q);
  bool setLowerBound(const std::vector<OSQPFloat>& l);
  bool setUpperBound(const std::vector<OSQPFloat>& u);

  void setVerbosity(bool verbosity);
  void setWarmStart(bool warmStart);
  void setAbsoluteTolerance(OSQPFloat tolerance);
  void setRelativeTolerance(OSQPFloat tolerance);
  void setMaxIteration(int maxIteration);
  ocp_b200_settings& settings() { return settings_; }

  bool initSolver();
  void printSolverData();
  bool solve();

  std::vector<OSQPFloat> getSolution();
  std::vector<OSQPFloat> getDualSolution();
  casadi::DM getSolutionAsDM();
  const std::vector<double>& getInfo() const { return info_; }

  void setSystem(casadi::DMVector localSystem);

 private:
  void clearSolver();
  std::vector<OSQPFloat> upperBound, lowerBound, gradient;
  casadi::DM hessianMatrix, linearConstraintMatrix;
  std::vector<OSQPFloat> solution_, dual_;
  std::vector<double> info_;
  ocp_b200_solver* solver_;
  ocp_b200_settings settings_;
  int numOfVariables_;
  int numOfConstraints_;
  bool isInitialized_;
  bool verbose_;
};
