// Benchmark / test problems for the CUDA_SQP path.
//
// The reference ships no concrete OptimalControlProblem subclass and no H=20 test OCP
// (SURVEY.md "Reading notes"): deployConstraintsAndAddCost() is pure virtual
// (include/optimal_control_problem/OptimalControlProblem.h:101) and test/test.cpp only
// holds 8 small NLPs.  The subclasses here are the user-side code a caller of the
// reference would write, authored to the shapes BASELINE.json names:
//   quadrotor   nx=12 nu=4   dt=0.005 H=20   (configs[0..2])
//   centroidal  nx=24 nu=12  dt=0.01  H=50   friction pyramids (configs[3])
//   cartpole    nx=4  nu=1   dt=0.01  H=200  (configs[4])
// plus the 8 NLPs of test/test.cpp:13-211 as known-answer cases.
#pragma once

#include <memory>
#include <string>
#include <vector>

#include "optimal_control_problem/OptimalControlProblem.h"

namespace ocp_problems {

// YAML text of the `optimal_control_problem:` node for a named problem.  horizon <= 0 keeps
// the problem's own horizon; alpha/step_num are SQP_settings.
std::string default_yaml(const std::string& name, int horizon, double alpha, int step_num, bool verbose);

// Builds the problem (constructor + setReference + deployConstraintsAndAddCost); the caller
// decides whether to genSolver() (GPU) or to hand getNlp() to the oracle.
std::unique_ptr<OptimalControlProblem> make_problem(const std::string& name, const std::string& yaml_text);

int state_size(const std::string& name);  // nx (= |p| for all three problems)

// Deterministic synthetic inputs: B first frames [B*nf] and references [B*np]
// drawn with splitmix64 from the per-problem distributions.
void sample_inputs(const std::string& name, int B, unsigned long long seed, std::vector<double>& frames,
                   std::vector<double>& refs);

// test/test.cpp cases 1..8: nlp, solver arguments and the analytic optimum (empty for
// case 8, whose QP is non-convex).
struct KatCase {
  casadi::SXDict nlp;
  casadi::DMDict arg;
  std::vector<double> expected;
  std::string description;
};
KatCase make_kat(int id);

}  // namespace ocp_problems
