// OptimalControlProblem: the user-facing front-end of the CUDA_SQP path.
// Public surface of the reference class (include/optimal_control_problem/
// OptimalControlProblem.h:13-107): YAML::Node constructor, cost / constraint /
// reference registration, genSolver(), computeOptimalTrajectory().  Users derive
// from it and implement deployConstraintsAndAddCost().
//
// Scope: only solve_method CUDA_SQP is built.  IPOPT, SQP and
// MIXED are CasADi nlpsol plugins outside the named hot path; selecting them is
// rejected at construction time.  The ROS 2 package lookup
// (ament_index_cpp::get_package_share_directory, OptimalControlProblem.cpp:18) is
// replaced by $OCP_B200_SHARE_DIR or ./ocp_b200_share.
#pragma once

#include <iostream>
#include <memory>
#include <string>
#include <vector>

#include "casadi/casadi.hpp"
#include "optimal_control_problem/OCP_config/OCPConfig.h"
#include "optimal_control_problem/sqp_solver/SQPOptimizationSolver.h"
#include "yaml-cpp/yaml.h"

class OptimalControlProblem {
 private:
  class SolverSettings {
   public:
    enum class SolverType { IPOPT, SQP, CUDA_SQP, MIXED };
    struct SQPSettings {
      double alpha{0.1};
      int stepNum{10};
    };
    bool verbose{true};
    bool genCode{false};
    bool recompile{false};
    bool loadLib{false};
    bool warmStart{true};
    int maxIter{1000};
    SolverType solverType{SolverType::CUDA_SQP};
    SQPSettings SQP_settings;
  };

  YAML::Node configNode_;
  SolverSettings solverSettings;

  std::vector<casadi::SX> constraints_;
  std::vector<std::string> constraintNames_;
  std::vector<casadi::DM> constraintLowerBounds_;
  std::vector<casadi::DM> constraintUpperBounds_;
  casadi::SXVector costs_;
  bool setInitialGuess_{false};
  bool firstTime_{true};
  casadi::DM optimalTrajectory_;
  std::string packagePath_;
  std::string problemName_{"ocp"};
  std::shared_ptr<SQPOptimizationSolver> OSQPSolverPtr_;

  bool validateConfig(const YAML::Node& config);
  bool checkDirectoryPermissions(const std::string& path);

 public:
  typedef SolverSettings::SolverType SolverType;
  std::unique_ptr<OCPConfig> OCPConfigPtr_;

  void setSolverType(SolverSettings::SolverType type);
  SolverSettings::SolverType getSolverType() const;

  casadi::SX getReference() const;
  casadi::DM getOptimalTrajectory();

  casadi::SX reference_;
  casadi::SX totalCost_;

  void genSolver();
  void computeOptimalTrajectory(const casadi::DM& frame, const casadi::DM& reference);
  // Batched sibling: frames [B*frameSize], references [B*|p|]; warm-starts every
  // instance from its previous optimum.  Returns B trajectories, instance-major.
  const std::vector<double>& computeOptimalTrajectoryBatch(int B, const std::vector<double>& frames,
                                                           const std::vector<double>& references);
  const std::vector<double>& getBatchObjectives() const { return batchObjective_; }
  const std::vector<double>& getBatchStats() const { return batchStats_; }
  void resetWarmStart();
  void shiftBatchTrajectory();   // MPC tick: move every stored batch iterate one stage forward

  void setReference(const casadi::SX& reference);
  void setProblemName(const std::string& name) { problemName_ = name; }

  explicit OptimalControlProblem(YAML::Node);
  virtual ~OptimalControlProblem() = default;

  void addScalarCost(const casadi::SX& cost);
  void addVectorCost(const casadi::DM& param, const casadi::SX& cost);
  void addVectorCost(const std::vector<double>& param, const casadi::SX& cost);

  void addInequalityConstraint(const std::string& constraintName, const casadi::DM& lowerBound,
                               const casadi::SX& expression, const casadi::DM& upperBound);
  void addEquationConstraint(const std::string& constraintName, const casadi::SX& leftSX,
                             const casadi::SX& rightSX);
  void addEquationConstraint(const std::string& constraintName, const casadi::SX& expression);

  casadi::SX getCostFunction();
  casadi::DMVector getConstraintLowerBounds() const;
  casadi::DMVector getConstraintUpperBounds() const;
  std::vector<casadi::SX> getConstraints() const;
  const std::vector<std::string>& getConstraintNames() const { return constraintNames_; }

  virtual void deployConstraintsAndAddCost() = 0;

  bool solverInputCheck(std::map<std::string, casadi::DM> arg) const;

  // nlp {x, f, g, p} exactly as genSolver() hands it to the SQP driver
  // (reference OptimalControlProblem.cpp:235-240)
  casadi::SXDict getNlp();
  // solver arguments {lbx, ubx, lbg, ubg, x0, p} as computeOptimalTrajectory() builds
  // them (reference OptimalControlProblem.cpp:91-114)
  casadi::DMDict buildSolverArguments(const casadi::DM& frame, const casadi::DM& reference);
  std::shared_ptr<SQPOptimizationSolver> getSolver() { return OSQPSolverPtr_; }
  double getSqpAlpha() const { return solverSettings.SQP_settings.alpha; }
  int getSqpStepNum() const { return solverSettings.SQP_settings.stepNum; }
  bool getVerbose() const { return solverSettings.verbose; }

 private:
  std::vector<double> batchTrajectory_, batchObjective_, batchStats_;
  int batchSize_{0};
};

std::ostream& operator<<(std::ostream& os, const OptimalControlProblem& ocp);
