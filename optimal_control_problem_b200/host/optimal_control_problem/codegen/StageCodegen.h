// StageCodegen: turns the SQP local system (SQPOptimizationSolver.cpp:58-77) into
// CUDA source -- one straight-line __device__ function per stage template,
// evaluated by one warp per (instance, stage group) -- and compiles it with nvcc
// for sm_100a into a library implementing include/ocp_b200_model.h.
//
// This takes the place of the reference's run-time code generation hook
// (casadi codegen + std::system("gcc ...") in OptimalControlProblem.cpp:602-640,
// which only exists for the IPOPT/SQP plugins) on the CUDA_SQP path.
#pragma once

#include <string>

#include "casadi/casadi.hpp"

namespace ocp_codegen {

struct ModelSpec {
  std::string name;
  casadi::SX p, x;        // symbolic parameters and decision variables (dense columns)
  casadi::SX f, g;        // objective (1x1) and constraints (ng x 1, may be empty)
  casadi::SX grad;        // gradient of f w.r.t. w = [p; x]  (n x 1, dense)
  casadi::SX hess;        // Hessian of f w.r.t. w            (n x n, structural pattern)
  casadi::SX jac;         // Jacobian of c = [p; x; g] w.r.t. w (m x n, structural pattern)
  int nf{0}, horizon{1};  // stage layout of x: horizon frames of nf variables
};

struct ModelSource {
  std::string source;     // complete .cu translation unit
  unsigned long long hash{0};
  int num_groups{0}, num_templates{0};
  size_t num_statements{0};
};

ModelSource generate(const ModelSpec& spec);

// Writes <code_dir>/<name>_<hash>.cu and compiles it to <code_dir>/<name>_<hash>.so unless that
// file already exists.  Returns the library path.  Throws std::runtime_error when nvcc fails.
std::string compile(const ModelSource& src, const std::string& name, const std::string& code_dir,
                    bool verbose);

// A CasADi-generated C file (Function::generate / CodeGenerator layout) as the stage library: `local_system_fn`
// maps (p, x, l, u) -> (H, grad f, J, l - c, u - c), `objective_fn` maps (p, x) -> f; nf / horizon give the stage
// layout of x.  Compiles <code_dir>/<name>_casadi_<hash>.so (host build for the sparsity queries, then nvcc with the
// C functions as __device__ code, one thread per instance) and returns its path.  See CasadiCInterop.cpp.
std::string compile_casadi_c(const std::string& c_file, const std::string& local_system_fn, const std::string& objective_fn,
                             const std::string& name, int nf, int horizon, const std::string& code_dir, bool verbose);

// directory holding ocp_b200.h / ocp_b200_model.h ($OCP_B200_INCLUDE_DIR or relative to this file)
std::string include_dir();

}  // namespace ocp_codegen
