/* ocp_b200_model.h -- ABI of a compiled stage-function library.
 *
 * The reference evaluates the per-iteration QP data by interpreting one
 * whole-horizon casadi::Function on the host (localSystemFunction_,
 * src/sqp_solver/SQPOptimizationSolver.cpp:74-77, 116-117).  Here genSolver()
 * emits the same quantities as straight-line CUDA device functions, one per
 * stage template, compiles them with nvcc (sm_100a) into a shared library and
 * hands the library path to ocp_b200_create().  The library exports exactly the
 * symbols below; every pointer is a device pointer.
 */
#ifndef OCP_B200_MODEL_H
#define OCP_B200_MODEL_H

#ifdef __cplusplus
extern "C" {
#endif

#define OCP_B200_MODEL_ABI_VERSION 1

typedef struct ocp_b200_model_info {
  int abi_version;
  int np, nf, horizon, ng;
  int n, m;                 /* n = np + nf*horizon, m = n + ng                   */
  int nnz_h, nnz_a;
  const int* h_colptr;      /* host arrays, n+1 / nnz_h (full symmetric pattern) */
  const int* h_rowidx;
  const int* a_colptr;      /* host arrays, n+1 / nnz_a                          */
  const int* a_rowidx;
  int num_groups;           /* warps of work per instance in assemble            */
  int num_templates;        /* distinct stage programs after de-duplication      */
  const char* name;
  unsigned long long source_hash;
} ocp_b200_model_info;

const ocp_b200_model_info* ocp_b200_model_get_info(void);

/* Local system at (p, x) for B instances: one warp per (instance, stage group).
 *   x [B*N] p [B*np] frames [B*nf or NULL]; lbx,ubx [N], lbg,ubg [ng] shared
 *   h_vals [B*ld_h] q [B*ld_n] a_vals [B*ld_a] l,u [B*ld_m]   (ld_* = row pitch in doubles)
 * Writes H, grad f, J and l - c, u - c with c = [p; x; g]
 * (SQPOptimizationSolver.cpp:58-71). */
int ocp_b200_model_assemble(int B, const double* x, const double* p, const double* frames,
                            const double* lbx, const double* ubx, const double* lbg, const double* ubg,
                            double* h_vals, int ld_h, double* q, int ld_n, double* a_vals, int ld_a,
                            double* l, double* u, int ld_m, void* stream);

/* f(p, x) for B instances (SQPOptimizationSolver.cpp:180-181) */
int ocp_b200_model_objective(int B, const double* x, const double* p, double* f, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* OCP_B200_MODEL_H */
