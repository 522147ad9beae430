/* ocp_b200.h -- C ABI of the B200-native CUDA_SQP solve path.
 *
 * This is the drop-in boundary.  The reference
 * (LockedFlysher/optimal_control_problem) has no FFI of its own: its CUDA_SQP path is a
 * stack of C++ classes ending in the OSQP v1 C API (osqp_setup / osqp_solve /
 * osqp_cleanup, reached through OsqpEigen at src/sqp_solver/CuCaQP.cpp:183-224).
 * The entry points below are what those classes bind instead; each one names the
 * reference interface it replaces.  Plain C: pointers + sizes, `int` status
 * returns (0 = OCP_B200_OK), no exceptions cross the ABI, the caller owns every
 * host buffer, the library owns every device buffer.  One handle per (host
 * thread, CUDA device); handles are not re-entrant.
 *
 * All floating point data is FP64.  All index arrays are int32, CSC
 * (column-compressed, strictly increasing row indices per column) -- the CCS
 * order CasADi produces and CuCaQP.h:105-137 preserves.
 *
 * There is no CPU fallback behind this ABI: if no CUDA device (or no compiled
 * stage library) is available the calls fail with an error code.
 */
#ifndef OCP_B200_H
#define OCP_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define OCP_B200_ABI_VERSION 1

/* ---- status codes ------------------------------------------------------ */
#define OCP_B200_OK               0
#define OCP_B200_ERR_INVALID      1   /* bad argument / inconsistent sizes        */
#define OCP_B200_ERR_CUDA         2   /* CUDA runtime error (see last_error)      */
#define OCP_B200_ERR_NO_DEVICE    3   /* no CUDA device: there is no CPU fallback */
#define OCP_B200_ERR_MODEL        4   /* stage library missing / ABI mismatch     */
#define OCP_B200_ERR_UNSUPPORTED  5   /* problem shape not supported by this build*/

/* ---- per-QP status (OSQP v1 numbering, osqp_api_constants.h) ------------ */
#define OCP_B200_QP_SOLVED               1
#define OCP_B200_QP_SOLVED_INACCURATE    2
#define OCP_B200_QP_PRIMAL_INFEASIBLE    3
#define OCP_B200_QP_DUAL_INFEASIBLE      5
#define OCP_B200_QP_MAX_ITER_REACHED     7
#define OCP_B200_QP_UNSOLVED            11

/* ---- reduced-KKT linear system (K = P + sigma I + A' diag(rho) A) ------- */
#define OCP_B200_PRECOND_DIAGONAL      0  /* PCG, Jacobi: what OSQP's cuda backend ships       */
#define OCP_B200_PRECOND_BLOCK_JACOBI  1  /* PCG, one dense block per stage (+ one for p)      */
#define OCP_B200_PRECOND_BLOCK_TRIDIAG 2  /* stage-block tridiagonal LDL' with a dense border
                                           * for p.  K of a multiple-shooting OCP has exactly
                                           * this structure, so the factorisation IS K and one
                                           * application solves the system (no CG iterations).
                                           * Patterns without that structure fall back to
                                           * BLOCK_JACOBI PCG (see ocp_b200_get_dims).        */

/* Settings.  Defaults (ocp_b200_default_settings) are the values the reference
 * runs with: SQPOptimizationSolver.cpp:81-85 (eps 1e-3, max_iter 10000), the OSQP
 * v1.0.0.beta1 defaults for everything it leaves untouched, and
 * OptimalControlProblem.h:24-27 (alpha 0.1, step_num 10). */
typedef struct ocp_b200_settings {
  /* SQP driver: SQPOptimizationSolver.cpp:15-16, 137, 171-177 */
  double sqp_alpha;              /* fixed step length ("SQP_step")            */
  int    sqp_step_num;           /* exact number of SQP steps ("ADMM_step")   */
  /* ADMM (OSQP) */
  double eps_abs, eps_rel;       /* 1e-3, 1e-3                                */
  double eps_prim_inf, eps_dual_inf; /* 1e-4, 1e-4                            */
  int    admm_max_iter;          /* 10000                                     */
  double rho;                    /* 0.1                                       */
  double sigma;                  /* 1e-6                                      */
  double relax;                  /* OSQP "alpha", 1.6                         */
  int    scaling_iters;          /* Ruiz passes, 10                           */
  int    check_termination;      /* 25                                        */
  int    adaptive_rho;           /* 1                                         */
  int    adaptive_rho_interval;  /* 0 = OSQP's deterministic rule: 4*check_termination */
  double adaptive_rho_tolerance; /* 5                                         */
  /* reduced-KKT PCG */
  int    pcg_max_iter;           /* per ADMM iteration                        */
  double pcg_tol;                /* stop when ||r||_2 <= pcg_tol * ||rhs||_2  */
  int    pcg_precond;            /* OCP_B200_PRECOND_*                        */
} ocp_b200_settings;

/* Problem description.  n = np + nf*horizon, m = n + ng
 * (SQPOptimizationSolver.cpp:50, 54, 80: w = [p; x], c = [p; x; g]). */
typedef struct ocp_b200_problem_desc {
  int np;        /* |p|, reference parameters                                  */
  int nf;        /* frame size (OCPConfig::getFrameSize)                       */
  int horizon;   /* OCPConfig::getHorizon                                      */
  int ng;        /* |g|, user constraints                                      */
  /* Hessian of the objective w.r.t. w, FULL symmetric pattern, n-by-n, exactly
   * what localSystemFunction output 0 has.  Like OsqpEigen the solver reads
   * only the upper triangle and mirrors it. */
  int nnz_h; const int* h_colptr; const int* h_rowidx;
  /* Jacobian of c w.r.t. w, m-by-n (localSystemFunction output 2) */
  int nnz_a; const int* a_colptr; const int* a_rowidx;
  /* nvcc-compiled stage-function library implementing ocp_b200_model.h; NULL
   * makes a QP-only handle (ocp_b200_qp_* entry points). */
  const char* model_library;
  /* optional preconditioner block partition of the n columns: num_blocks+1
   * ascending offsets.  NULL: [p | frame 0 | ... | frame H-1]. */
  int num_blocks; const int* block_ptr;
  int device;    /* CUDA device ordinal                                        */
} ocp_b200_problem_desc;

typedef struct ocp_b200_solver ocp_b200_solver;

/* per-instance statistics written by the SQP entry points (doubles) */
#define OCP_B200_NSTATS 12
#define OCP_B200_STAT_QP_STATUS    0  /* status of the last QP                 */
#define OCP_B200_STAT_SQP_STEPS    1
#define OCP_B200_STAT_ADMM_ITERS   2  /* summed over SQP steps                 */
#define OCP_B200_STAT_PCG_ITERS    3  /* summed over SQP steps                 */
#define OCP_B200_STAT_PRIM_RES     4  /* last QP, unscaled                     */
#define OCP_B200_STAT_DUAL_RES     5
#define OCP_B200_STAT_OBJECTIVE    6  /* f(p, x) after the last step           */
#define OCP_B200_STAT_RHO_UPDATES  7
#define OCP_B200_STAT_LAST_ADMM    8  /* ADMM iterations of the last QP        */
#define OCP_B200_STAT_LAST_RHO     9
#define OCP_B200_STAT_CHECKS      10  /* residual passes, summed               */
#define OCP_B200_STAT_STEP_NORM   11  /* ||alpha*dx||_2 of the last step       */

/* per-QP info written by the QP entry points (doubles) */
#define OCP_B200_NINFO 8
#define OCP_B200_INFO_STATUS       0
#define OCP_B200_INFO_ITERS        1
#define OCP_B200_INFO_PCG_ITERS    2
#define OCP_B200_INFO_PRIM_RES     3
#define OCP_B200_INFO_DUAL_RES     4
#define OCP_B200_INFO_RHO          5
#define OCP_B200_INFO_RHO_UPDATES  6
#define OCP_B200_INFO_CHECKS       7

/* replaces osqp_set_default_settings + SQPOptimizationSolver.cpp:81-85 */
void ocp_b200_default_settings(ocp_b200_settings* s);

/* replaces CuCaQP::setDimension + the symbolic half of osqp_setup
 * (CuCaQP.cpp:22-41, 183-197): fixes dimensions and sparsity once, uploads the
 * index structures, loads the stage library. */
int ocp_b200_create(const ocp_b200_problem_desc* desc, const ocp_b200_settings* settings,
                    ocp_b200_solver** out);
/* replaces osqp_cleanup (CuCaQP.cpp:16-21) */
int ocp_b200_destroy(ocp_b200_solver* s);
/* replaces CuCaQP::setAbsoluteTolerance / setRelativeTolerance / setMaxIteration ... (CuCaQP.cpp:163-181) */
int ocp_b200_update_settings(ocp_b200_solver* s, const ocp_b200_settings* settings);
int ocp_b200_get_settings(const ocp_b200_solver* s, ocp_b200_settings* out);

/* replaces SQPOptimizationSolver::getOptimalSolution (SQPOptimizationSolver.cpp:127-216)
 * for B independent instances.  Host buffers; H2D and D2H copies happen inside.
 *   frames [B*nf]  first frame of every instance, pinned into lbx/ubx[0:nf]
 *                  (OptimalControlProblem.cpp:93-96); NULL = use lbx/ubx as given
 *   p      [B*np]  reference parameters
 *   lbx,ubx[N]     variable bounds shared by the batch (N = nf*horizon)
 *   lbg,ubg[ng]    constraint bounds shared by the batch
 *   x_inout[B*N]   in: current iterate (the reference's persistent result_["x"]);
 *                  out: iterate after sqp_step_num steps
 *   f_out  [B]     objective at the returned iterate (may be NULL)
 *   stats  [B*OCP_B200_NSTATS] (may be NULL) */
int ocp_b200_solve_batch(ocp_b200_solver* s, int B, const double* frames, const double* p,
                         const double* lbx, const double* ubx, const double* lbg, const double* ubg,
                         double* x_inout, double* f_out, double* stats);

/* Same, with every pointer a DEVICE pointer on the solver's device; asynchronous
 * on `stream` (a cudaStream_t passed as void*). */
int ocp_b200_solve_batch_device(ocp_b200_solver* s, int B, const double* d_frames, const double* d_p,
                                const double* d_lbx, const double* d_ubx, const double* d_lbg,
                                const double* d_ubg, double* d_x_inout, double* d_f_out, double* d_stats,
                                void* stream);

/* MPC tick loop.  The reference keeps its iterate across calls
 * (result_["x"], SQPOptimizationSolver.cpp:215) exactly as the last solve left it; a
 * receding-horizon caller shifts it by one stage before the next tick: frame k <- frame k+1
 * for k < horizon-1, the last frame is repeated.  In place on d_x [B*N] (device pointer),
 * asynchronous on `stream`. */
int ocp_b200_shift_iterate_device(ocp_b200_solver* s, int B, double* d_x, void* stream);

/* Parity hook for SQPOptimizationSolver::getLocalSystem (SQPOptimizationSolver.cpp:100-120):
 * evaluates the local system at x for B instances and returns it, values in CCS
 * order of the patterns given at create time.
 *   h_vals[B*nnz_h] q[B*n] a_vals[B*nnz_a] l[B*m] u[B*m]  (host) */
int ocp_b200_export_qp(ocp_b200_solver* s, int B, const double* frames, const double* p,
                       const double* lbx, const double* ubx, const double* lbg, const double* ubg,
                       const double* x, double* h_vals, double* q, double* a_vals, double* l, double* u);

/* replaces CuCaQP::setSystem + initSolver + solve + getSolution
 * (CuCaQP.cpp:271-288, 183-224): B independent QPs sharing the create-time
 * patterns, cold-started (x = z = y = 0) like the reference does every SQP step.
 *   x_out[B*n] primal, y_out[B*m] dual (may be NULL), info[B*OCP_B200_NINFO] (may be NULL) */
int ocp_b200_qp_solve_batch(ocp_b200_solver* s, int B, const double* h_vals, const double* q,
                            const double* a_vals, const double* l, const double* u, double* x_out,
                            double* y_out, double* info);

/* Parity hook for osqp_solve's iteration history: solves ONE QP and records, at
 * every termination check, [iter, prim_res, dual_res, rho, pcg_iters_so_far, status]. */
#define OCP_B200_TRACE_WIDTH 6
int ocp_b200_admm_trace(ocp_b200_solver* s, const double* h_vals, const double* q, const double* a_vals,
                        const double* l, const double* u, int max_records, double* trace,
                        int* n_records, double* x_out, double* y_out);

/* counters for bench.py: kernels launched by this handle since creation */
long long ocp_b200_launch_count(const ocp_b200_solver* s);
/* Optional per-kernel device timing for bench.py's roofline: when enabled, every kernel launch
 * of this handle is bracketed by a CUDA event pair on the launching stream.  get_profile waits
 * for the recorded events and returns accumulated milliseconds and launch counts per kind.
 * enabled = 1: event pairs only; 2: additionally the per-phase cycle counters below. */
#define OCP_B200_NPROF 3
#define OCP_B200_PROF_ADMM      0  /* admm_solve_kernel (one launch per SQP step)   */
#define OCP_B200_PROF_ASSEMBLE  1  /* stage-library assembly kernel                 */
#define OCP_B200_PROF_OBJECTIVE 2  /* objective + stats store                       */
int ocp_b200_set_profiling(ocp_b200_solver* s, int enabled);
/* While profiling is on, CTA 0 of the direct kernel also accumulates SM cycles per phase of the
 * QP solves it runs (clock64 on one thread); get_phase_cycles returns and clears them. */
#define OCP_B200_NPHASE 14
#define OCP_B200_PHASE_LOAD         0
#define OCP_B200_PHASE_SCALE        1  /* Ruiz equilibration                        */
#define OCP_B200_PHASE_KKT_ASSEMBLE 2  /* K = P + sigma I + A' rho A into blocks    */
#define OCP_B200_PHASE_FACTOR       3  /* block-tridiagonal LDL'                    */
#define OCP_B200_PHASE_RHS          4
#define OCP_B200_PHASE_SOLVE        5
#define OCP_B200_PHASE_UPDATE       6  /* z~ = A x~, relaxation, projection, duals  */
#define OCP_B200_PHASE_CHECK        7  /* residuals, termination, rho adaptation    */
/* finer split of PHASE_SOLVE (counted in addition to it) */
#define OCP_B200_PHASE_SOLVE_FWD    8
#define OCP_B200_PHASE_SOLVE_BORDER 9
#define OCP_B200_PHASE_SOLVE_DIAG  10
#define OCP_B200_PHASE_SOLVE_BWD   11
/* finer split of PHASE_FACTOR (top chain of the twisted factorisation) */
#define OCP_B200_PHASE_FACTOR_INVERT 12
#define OCP_B200_PHASE_FACTOR_STEP   13
int ocp_b200_get_phase_cycles(ocp_b200_solver* s, long long* cycles);
int ocp_b200_get_profile(ocp_b200_solver* s, double* ms, long long* count, int reset);
/* dimensions of a handle: n, m, nnz_h, nnz_a, dynamic shared memory bytes per CTA, and
 * resident: bit 0 = all per-instance state is in shared memory, bit 1 = the direct
 * block-tridiagonal kernel is in use (0 = PCG kernel) */
int ocp_b200_get_dims(const ocp_b200_solver* s, int* n, int* m, int* nnz_h, int* nnz_a,
                      int* smem_bytes, int* resident);

/* ---- several GPUs of one node from ONE process ---------------------------------------------
 * The reference is single-instance, single-device; a C++ user of OptimalControlProblem reaches more than one GPU
 * through these entry points (OptimalControlProblem::computeOptimalTrajectoryBatch uses them when more than one
 * device is configured).  Instances are independent, so the batch is cut into contiguous blocks of ceil(B / ndev)
 * instances, one block per device in the order of `devices`; every device has its own handle, stream and slabs and
 * is driven by its own host thread; there is no traffic between the devices -- the "gather" is each device's
 * device-to-host copy landing in its slice of the caller's arrays.  (Device-resident multi-GPU batches, one process
 * per GPU with an NCCL all-gather, are what optimal_control_problem_b200/sharding.py and bench.py do.) */
typedef struct ocp_b200_multi ocp_b200_multi;

/* one handle per entry of devices[0..ndev) (a device may be listed more than once); desc->device is ignored */
int ocp_b200_create_multi(const ocp_b200_problem_desc* desc, const ocp_b200_settings* settings, const int* devices,
                          int ndev, ocp_b200_multi** out);
int ocp_b200_destroy_multi(ocp_b200_multi* m);
int ocp_b200_multi_update_settings(ocp_b200_multi* m, const ocp_b200_settings* settings);
/* offsets[0..ndev]: instance range [offsets[k], offsets[k+1]) goes to device k.  Pure arithmetic, needs no GPU. */
int ocp_b200_multi_partition(int B, int ndev, int* offsets);
int ocp_b200_multi_device_count(const ocp_b200_multi* m);
ocp_b200_solver* ocp_b200_multi_handle(ocp_b200_multi* m, int k);   /* borrowed: settings queries, profiling */
/* same arguments and results as ocp_b200_solve_batch; bit-identical to it for every instance */
int ocp_b200_solve_batch_multi(ocp_b200_multi* m, int B, const double* frames, const double* p, const double* lbx,
                               const double* ubx, const double* lbg, const double* ubg, double* x_inout,
                               double* f_out, double* stats);

/* launch plan and linear-system structure of a handle (diagnostics, roofline accounting): fills up to
 * `count` ints of v in the order of the OCP_B200_PLAN_* slots */
#define OCP_B200_PLAN_WIDE_PLACE      0   /* throughput plan: placement id (0 mixed, 1 all shared, 2 multi-CTA, 3 big) */
#define OCP_B200_PLAN_WIDE_THREADS    1
#define OCP_B200_PLAN_WIDE_SMEM       2   /* dynamic shared memory bytes per CTA */
#define OCP_B200_PLAN_WIDE_CTAS_SM    3   /* resident CTAs per SM */
#define OCP_B200_PLAN_WIDE_SLAB_KB    4   /* per-CTA global slab, KiB */
#define OCP_B200_PLAN_DEEP_PLACE      5   /* latency plan (batches of at most one instance per SM) */
#define OCP_B200_PLAN_DEEP_THREADS    6
#define OCP_B200_PLAN_DEEP_SMEM       7
#define OCP_B200_PLAN_DEEP_CTAS_SM    8
#define OCP_B200_PLAN_DEEP_SLAB_KB    9
#define OCP_B200_PLAN_TRI_OK         10   /* 1: K is bordered block tridiagonal and the direct kernel is used */
#define OCP_B200_PLAN_TRI_NP         11   /* border columns */
#define OCP_B200_PLAN_TRI_BS         12   /* block size */
#define OCP_B200_PLAN_TRI_NB         13   /* number of diagonal blocks */
#define OCP_B200_PLAN_NNZ_P          14   /* entries of the symmetrised upper triangle of H */
#define OCP_B200_PLAN_NUM_SMS        15
#define OCP_B200_PLAN_COUNT          16
int ocp_b200_get_plan(const ocp_b200_solver* s, int* v, int count);

const char* ocp_b200_last_error(void);
int ocp_b200_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* OCP_B200_H */
