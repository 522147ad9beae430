// SQPOptimizationSolver: fixed-step SQP driver of the CUDA_SQP path.
// Public surface of the reference class (include/optimal_control_problem/sqp_solver/
// SQPOptimizationSolver.h:9-81).  The constructor does the same symbolic work
// (src/sqp_solver/SQPOptimizationSolver.cpp:12-92); the solve loop
// (SQPOptimizationSolver.cpp:127-216) runs on the GPU through include/ocp_b200.h.
#pragma once

#include <memory>
#include <string>
#include <vector>

#include "casadi/casadi.hpp"
#include "ocp_b200.h"
#include "optimal_control_problem/sqp_solver/AutoDifferentiator.h"
#include "optimal_control_problem/sqp_solver/CuCaQP.h"

class SQPOptimizationSolver {
 public:
  // nlp: {"x","f"} required, {"g","p"} optional.  options: "max_iter" (= number of SQP
  // steps), "alpha", "verbose"; optional "code_dir" (where stage libraries are cached),
  // "name", "nf"/"horizon" (stage layout of x; default: one stage).
  explicit SQPOptimizationSolver(casadi::SXDict& nlp, casadi::Dict& options);
  ~SQPOptimizationSolver();

  // arg: lbx, ubx, lbg, ubg, p (optional) -> {x, f}.  x warm-starts from the previous call.
  casadi::DMDict getOptimalSolution(const casadi::DMDict& arg);

  // Batched sibling (north_star subsystem 3): B instances that differ in their first
  // frame, reference p and current iterate.  frames may be empty (bounds used as given).
  // x_inout [B*N] is read and overwritten; f_out [B]; stats [B*OCP_B200_NSTATS] optional.
  void getOptimalSolutionBatch(int B, const std::vector<double>& frames, const std::vector<double>& p,
                               const casadi::DM& lbx, const casadi::DM& ubx, const casadi::DM& lbg,
                               const casadi::DM& ubg, std::vector<double>& x_inout,
                               std::vector<double>& f_out, std::vector<double>* stats = nullptr);

  casadi::Function getSXLocalSystemFunction() const;
  casadi::Function getObjectiveFunction() const { return objectiveFunction_; }
  void setVerbose(bool verbose);

  // local system at the current iterate via the GPU assembly kernel (parity hook)
  casadi::DMVector getLocalSystemGPU(const casadi::DMDict& arg);

  // Batched solves on several GPUs of the node (include/ocp_b200.h: ocp_b200_create_multi): the batch is cut into
  // contiguous blocks, one per listed device.  An empty list / one entry = the single device of `options["device"]`.
  // The environment variable OCP_B200_DEVICES ("0,1,2,3") sets the list for programs that cannot be changed.
  void setDevices(const std::vector<int>& devices);
  const std::vector<int>& devices() const { return devices_; }

  void ensureDevice();  // creates the device solver; throws when no CUDA device is usable
  ocp_b200_solver* handle() { ensureDevice(); return handle_; }
  ocp_b200_settings& settings() { return settings_; }
  void applySettings();
  const std::string& modelLibrary() const { return modelLibrary_; }
  int numVariables() const { return n_; }
  int numConstraints() const { return m_; }
  int numParameters() const { return np_; }
  void resetIterate();
  // SQP schedule of the next solve: number of steps and step length (reference: fixed at construction)
  void setSchedule(int stepNum, double alpha);   // also pushed to the device handle when it exists
  double lastObjective() const { return densify(result_.at("f")).nonzeros().at(0); }

 private:
  std::shared_ptr<AutoDifferentiator> objectiveFunctionAutoDifferentiatorPtr_;
  std::shared_ptr<AutoDifferentiator> constraintsAutoDifferentiator_;

  int stepNum_;
  double alpha_;
  bool verbose_;
  casadi::DMDict result_;
  casadi::Function objectiveFunction_;
  // inputs [p, x, l, u] -> outputs [H, grad, A, l', u']  (reference SQPOptimizationSolver.h:66-72)
  casadi::Function localSystemFunction_;

  int np_{0}, N_{0}, ng_{0}, n_{0}, m_{0}, nf_{0}, horizon_{1}, device_{0};
  std::vector<int> hColptr_, hRowidx_, aColptr_, aRowidx_;
  std::string modelLibrary_;
  ocp_b200_solver* handle_{nullptr};
  ocp_b200_multi* multi_{nullptr};
  std::vector<int> devices_;
  ocp_b200_settings settings_;
};
