/* b2t.h -- C ABI of the B200-native batched SQP / Schur-complement / GBD-PCG trajectory-optimisation path.
 *
 * One shared library is built per robot (topology baked in at code-generation time): libb2t_<robot>.so.
 * Every library exports exactly the symbols below.  All functions return 0 on success, a negative b2t_status
 * otherwise; none throws, aborts or prints.  The caller owns every input/output buffer; the library owns only
 * the opaque solver handle (its device workspace).  Thread-safe per handle.
 *
 * The reference (VCA-EPFL/TrajoptMPCReference) has no FFI: its boundary is the Python API.  Each entry point
 * names the reference interface it replaces (file:line under /root/reference):
 *
 *   b2t_solver_create        TrajoptMPCReference.__init__ (TrajoptMPCReference.py:31-43) + URDFPlant.__init__
 *                            (TrajoptPlant.py:274-281) + QuadraticCost/UrdfCost.__init__ (TrajoptCost.py:24-37,
 *                            373-396) + TrajoptConstraint.set_*_limits (TrajoptConstraint.py:191-208)
 *   b2t_set_trajectory / b2t_set_goals / b2t_set_multipliers
 *                            arguments x, u of SQP() (:510), cost.xg, BoxConstraint mu/lambda/phi (:23-25)
 *   b2t_sqp_solve            TrajoptMPCReference.SQP (:510-760), batched over independent instances
 *   b2t_ilqr_solve           the iLQR solver the reference names but does not ship (README.md:15-17, :21-27)
 *   b2t_get_trajectory / b2t_get_status / b2t_get_trace / b2t_get_multipliers
 *                            the 6-tuple SQP returns (:760), self.trace (:555-569), mu/lambda/phi state
 *   b2t_sqp_solve_host       one call from host buffers to host buffers (what examples/exampleHelpers.py:80 does)
 *   b2t_stage_dynamics       TrajoptPlant.integrator(.., return_gradient) (TrajoptPlant.py:83-205)
 *   b2t_stage_kkt            formKKTSystemBlocks (:200-271) + Schur complement and preconditioner blocks
 *                            (solveKKTSystem_Schur :419-424, PCG.compute_preconditioner PCG.py:166-212)
 *   b2t_stage_pcg            PCG.solve (PCG.py:66-111, 214)
 *   b2t_stage_recover        dxu = invG (g - C^T l) (:449-452)
 *   b2t_stage_merit          totalCost (:296-310), totalHardConstraintViolation (:273-294), D (:635-648)
 *   b2t_fetch                the reference's saved_* lists (exampleHelpers.py:136-154), one array at a time
 *
 * Array layouts at the ABI are the reference's:  x is [batch][nx][N] (C order, i.e. numpy (nx, N) per
 * instance), u is [batch][nu][N-1], goals xg [batch][nx] (B2T_COST_URDF_EE: (x, y, vx, vy) in the first 4 of the nx slots).  double precision at the boundary.
 */
#ifndef B2T_H
#define B2T_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

#define B2T_ABI_VERSION 2

typedef enum {
  B2T_OK = 0,
  B2T_ERR_INVALID = -1,      /* bad argument (the reference print()s and exit()s) */
  B2T_ERR_CUDA = -2,         /* CUDA runtime error; see b2t_last_error */
  B2T_ERR_UNSUPPORTED = -3,  /* e.g. the end-effector cost on a 1-joint robot, integrator type 4 */
  B2T_ERR_NOMEM = -4,
  B2T_ERR_UNFINISHED = -5    /* the pass budget ran out with instances still active (results of the others are valid) */
} b2t_status;

typedef enum { B2T_COST_QUADRATIC = 0, B2T_COST_URDF_EE = 1 } b2t_cost_kind;
/* BoxConstraint modes (TrajoptConstraint.py:27-51).  QUADRATIC_PENALTY / AUGMENTED_LAGRANGIAN are soft (all methods, SQP and iLQR).
 * ACTIVE_SET is hard: violated bounds enter the KKT system as rows (TrajoptMPCReference.py:238-248); exact methods N / S only.
 * FULL_SET (singular KKT in the reference) and ADMM_PROJECTION (unimplemented there) are not offered. */
typedef enum { B2T_LIMIT_NONE = 0, B2T_LIMIT_QUADRATIC_PENALTY = 1, B2T_LIMIT_AUGMENTED_LAGRANGIAN = 2, B2T_LIMIT_ACTIVE_SET = 3 } b2t_limit_mode;
/* SQPSolverMethods (TrajoptMPCReference.py:13-18).  N (dense KKT backslash, :313-359) and S (Schur backslash, :430-436) are
 * exact solves of the same linear system; both run the block-tridiagonal factorisation of the Schur complement here. */
typedef enum { B2T_METHOD_N = 0, B2T_METHOD_S = 1, B2T_METHOD_PCG_J = 2, B2T_METHOD_PCG_BJ = 3, B2T_METHOD_PCG_SS = 4 } b2t_method;
typedef enum { B2T_F64 = 0, B2T_F32 = 1 } b2t_dtype;
/* limit types, index into the per-type arrays below */
enum { B2T_LIM_JOINT = 0, B2T_LIM_VELOCITY = 1, B2T_LIM_TORQUE = 2 };

typedef struct {
  int batch;              /* independent MPC instances */
  int knots;              /* N */
  int integrator_type;    /* 0 euler, 1 semi-implicit euler, 2 midpoint, 3 rk3 -- 2 / 3 with the reference's own arithmetic (TrajoptPlant.py:92-205) */
  int dtype;              /* b2t_dtype: arithmetic type of the whole path */
  double dt;
  double gravity;         /* options['gravity'], default -9.81 (TrajoptPlant.py:31) */
  int cost_kind;          /* b2t_cost_kind */
  int qf_start;           /* QF_start, -1 = None (TrajoptCost.py:40-47) */
  const double* Q;        /* [nx*nx] row-major, host; B2T_COST_URDF_EE: the 4 x 4 weights of (x, y, vx, vy) packed in the first 16 */
  const double* QF;       /* [nx*nx] */
  const double* R;        /* [nu*nu] */
  int limit_mode[3];      /* b2t_limit_mode per limit type */
  const double* lower;    /* [nx+nu] lower bounds of z = [q; qd; u] (read where the type's mode != NONE) */
  const double* upper;    /* [nx+nu] */
  double mu_init[3], mu_factor[3], mu_max[3], phi_init[3], phi_factor[3];   /* TrajoptConstraint.py:40-44 */
  int hess_mode;          /* UrdfCost.hess_mode (TrajoptCost.py:391-395): 0 Gauss-Newton J^T Q J, 1 exact Hessian (ABI version 2) */
} b2t_problem_desc;

typedef struct {           /* TrajoptMPCReference.set_default_options (:91-115) */
  double exit_tolerance_linSys;
  int max_iter_linSys;
  double exit_tolerance_SQP;
  int max_iter_SQP;
  double alpha_factor, alpha_min;
  double rho_factor, rho_min, rho_max, rho_init;
  double expected_reduction_min, expected_reduction_max;
  double exit_tolerance_soft;
  int max_iter_soft;
  double merit_mu;        /* the reference hard-codes mu = 10 (:546) */
} b2t_options;

/* per-instance result record written by b2t_get_status: [batch][B2T_STATUS_FIELDS] ints */
enum { B2T_ST_EXIT_SQP = 0, B2T_ST_EXIT_SOFT, B2T_ST_OUTER_ITER, B2T_ST_SQP_ITER, B2T_ST_TOTAL_QP, B2T_ST_TOTAL_PCG,
       B2T_ST_TOTAL_TRIALS, B2T_ST_TRACE_ROWS, B2T_STATUS_FIELDS };
/* per-instance scalar results written by b2t_get_scalars: [batch][B2T_SCALAR_FIELDS] doubles */
enum { B2T_SC_J = 0, B2T_SC_C, B2T_SC_MERIT, B2T_SC_RHO, B2T_SCALAR_FIELDS };
/* one trace row (b2t_get_trace): [batch][trace_cap][B2T_TRACE_FIELDS] doubles, same keys as self.trace (:555-569) */
enum { B2T_TR_OUTER = 0, B2T_TR_ITER, B2T_TR_LS_ITER, B2T_TR_ALPHA, B2T_TR_RHO, B2T_TR_J, B2T_TR_C, B2T_TR_MERIT, B2T_TR_D,
       B2T_TR_RATIO, B2T_TR_INNER, B2T_TR_SUCCESS, B2T_TRACE_FIELDS };

/* arrays b2t_fetch can return (knot-major, doubles): name -> shape per instance */
typedef enum {
  B2T_ARR_X = 0,        /* [N][nx]   current iterate */
  B2T_ARR_U,            /* [N][nu]   (row N-1 unused) */
  B2T_ARR_XKP1,         /* [N][nx]   integrator(x_k,u_k) (row N-1 unused) */
  B2T_ARR_DQDD,         /* [N][n*3n] forward_dynamics_gradient (integrator types 0 / 1; 2 / 3 evaluate it per stage: fetch B2T_ARR_AB) */
  B2T_ARR_GHAT,         /* [N][m*m]  inv(G_k + rho I) */
  B2T_ARR_G,            /* [N][m]    gradient g_k */
  B2T_ARR_SD,           /* [N][nx*nx] diagonal blocks of S */
  B2T_ARR_SO,           /* [N][nx*nx] S_{k,k-1} (row 0 unused) */
  B2T_ARR_PD,           /* [N][nx*nx] diagonal blocks of the preconditioner */
  B2T_ARR_GAMMA,        /* [N][nx] */
  B2T_ARR_L,            /* [N][nx]   multipliers from PCG */
  B2T_ARR_DZ,           /* [N][m]    step [dx_k; du_k] */
  B2T_ARR_XN,           /* [N][nx]   line-search trial point */
  B2T_ARR_UN,           /* [N][nu] */
  /* per-knot cost terms at the current (x, u), evaluated on demand: TrajoptCost.value / gradient / hessian (TrajoptCost.py:49-83, 402-519) */
  B2T_ARR_COST_VALUE,   /* [N][1] */
  B2T_ARR_COST_GRAD,    /* [N][m]   (control part zero at the terminal knot) */
  B2T_ARR_COST_HESS,    /* [N][m*m] */
  B2T_ARR_COST_ERR,     /* [N][nx]  state error of the cost's state map (UrdfCost.delta_x, TrajoptCost.py:425-435) */
  B2T_ARR_KKT_HESS,     /* [N][m*m] G_k = cost Hessian + gck gck^T, without rho (TrajoptMPCReference.py:214-224) */
  B2T_ARR_AB,           /* [N][nx*m] [A_k B_k] of the integrator, from the last dynamics pass (row N-1 zero) (:229-233) */
  B2T_ARR_SOFT_VALUE,   /* [N][1]   value_soft_constraints at (x_k, u_k) with the current multipliers (TrajoptConstraint.py:295-308) */
  B2T_ARR_SOFT_GRAD,    /* [N][m]   summed penalty gradient gck (jacobian_soft_constraints, :310-340; element-wise restatement) */
  B2T_ARR_NU_TRACE,     /* [128]    per instance: |r^T Pinv r| of PCG iteration 0..127 of the last b2t_stage_pcg (PCG.py:82,95) */
  B2T_ARR_COST_JTOT,    /* [N][nx*nx] Jacobian of the cost's state map at x_k: UrdfCost jacobian_tot_state (ne x nx row-major in the slot; identity for
                           the joint-space cost) -- the entries of UrdfCost.saved_Jacobian_tot_state (TrajoptCost.py:439,480) */
  B2T_ARR_PLANT_TERMS   /* [N][2n + 3n*n] per knot [c | qdd | Minv (n x n) | d rnea / d(q, qd) (n x 2n)] at (x_k, u_k): the entries of
                           URDFPlant.saved_c / saved_qdd / saved_Minv / saved_dc_du (TrajoptPlant.py:283-323); row N-1 zero */
} b2t_array;

typedef struct b2t_solver b2t_solver;

int b2t_abi_version(void);
const char* b2t_model_name(void);
const char* b2t_model_digest(void);
int b2t_model_dims(int* nq, int* nx, int* nu);
const char* b2t_last_error(void);
void b2t_default_options(b2t_options* o);

int b2t_solver_create(const b2t_problem_desc* desc, int device, b2t_solver** out);
int b2t_solver_destroy(b2t_solver* s);
size_t b2t_workspace_bytes(const b2t_solver* s);
int b2t_set_stream(b2t_solver* s, void* cuda_stream);

/* inputs (host pointers unless on_device != 0, then device pointers of the same layout and type double) */
int b2t_set_trajectory(b2t_solver* s, const double* x, const double* u, int on_device);
int b2t_set_goals(b2t_solver* s, const double* xg, int on_device);
/* xs [batch][nx] host: initial-state target of the first constraint row c_0 = x_0 - xs; b2t_set_trajectory sets xs = x[:,0]
 * like SQP() does (TrajoptMPCReference.py:527) */
int b2t_set_initial_state(b2t_solver* s, const double* xs);
int b2t_set_multipliers(b2t_solver* s, const double* mu, const double* lam, const double* phi);   /* [batch][2m][N] host */
int b2t_reset_multipliers(b2t_solver* s);

int b2t_sqp_solve(b2t_solver* s, int method, const b2t_options* opts);
/* Recording hook (replaces the reference's saved_* lists, TrajoptMPCReference.py:46-70, filled inside SQP when examples run with
 * record=True, examples/exampleHelpers.py:85-154).  b2t_sqp_solve calls `hook(user, event, pass)` on the host between kernel
 * launches, stream idle: B2T_HOOK_LINSYS after the linear system of an SQP iteration is solved (KKT blocks, Ghat, S, gamma,
 * preconditioner, l, dz of every ACTIVE instance are valid and can be read with b2t_fetch; all blocks are written in this mode), and
 * B2T_HOOK_STEP after its line search and exit logic (x, u, status updated).  A non-zero return aborts the solve with
 * B2T_ERR_INVALID.  NULL removes the hook.  The hook must not call b2t_sqp_solve / b2t_stage_* on the same handle. */
typedef int (*b2t_iteration_hook)(void* user, int event, int pass);
enum { B2T_HOOK_LINSYS = 1, B2T_HOOK_STEP = 2 };
int b2t_set_iteration_hook(b2t_solver* s, b2t_iteration_hook hook, void* user);
/* iLQR / DDP on the same problem description (MPCSolverMethods.iLQR, TrajoptMPCReference.py:21-27; the reference ships no
 * implementation -- specification: oracle/ilqr.py).  x[:,0] is the start state, the state trajectory is re-rolled from u. */
int b2t_ilqr_solve(b2t_solver* s, const b2t_options* opts);

/* Receding-horizon step (MPC loop around SQP / iLQR): reports the current initial state x_0 and the control u_0
 * that is applied, takes the next initial state from `x_next` ([batch][nx] host; NULL = simulate x_0, u_0 with the plant's own
 * integrator), shifts x, u and the soft-constraint multipliers one knot to the left (TrajoptConstraint.shift_soft_constraint_constants,
 * TrajoptConstraint.py:168-176) and sets xs.  The workspace then holds the warm start of the next solve.  Outputs may be NULL. */
int b2t_mpc_shift(b2t_solver* s, const double* x_next, double* x0_out, double* u0_out, double* xnext_out);

int b2t_get_trajectory(b2t_solver* s, double* x, double* u, int on_device);
int b2t_get_status(b2t_solver* s, int* status);
int b2t_get_scalars(b2t_solver* s, double* scalars);
int b2t_get_trace(b2t_solver* s, double* trace, int trace_cap);
int b2t_get_multipliers(b2t_solver* s, double* mu, double* lam, double* phi);
/* number of kernels launched and seconds spent (CUDA events) by the last b2t_sqp_solve */
int b2t_get_launch_stats(b2t_solver* s, long long* launches, double* device_seconds);
/* CUDA-event time of each kernel family accumulated over the last solve: [B2T_KERNEL_FAMILIES] seconds and launch counts */
enum { B2T_K_FD = 0, B2T_K_FDGRAD, B2T_K_KKT, B2T_K_SCHUR, B2T_K_PCG, B2T_K_RECOVER, B2T_K_TRIAL, B2T_K_MERIT, B2T_K_CTRL, B2T_KERNEL_FAMILIES };
/* mode 0: no events; 1: every launch bracketed by CUDA events on the solver's stream; 2 + f: only the launches of family f */
int b2t_set_profiling(b2t_solver* s, int mode);
int b2t_get_kernel_times(b2t_solver* s, double* seconds, long long* launches);
/* SQP passes of the last b2t_sqp_solve (one pass = one SQP iteration of every instance still active: the batched form of the loop
 * TrajoptMPCReference.py:573-750) and the number of instances still active after each of the first `cap` passes */
int b2t_get_pass_trace(b2t_solver* s, int* active_counts, int cap, int* passes);
/* name of the PCG kernel b2t_sqp_solve launches for this problem ("k_pcg3", "k_pcg2", "k_pcg"): the key of the ncu captures in profiles/ */
const char* b2t_pcg_kernel_name(b2t_solver* s);

/* host -> device, solve, device -> host in one call (pinned staging inside the handle) */
int b2t_sqp_solve_host(b2t_solver* s, const double* x0, const double* u0, const double* xg, int method, const b2t_options* opts,
                       double* x_out, double* u_out, int* status_out);

/* single stages on the handle's current (x, u), all instances; results via b2t_fetch */
int b2t_stage_dynamics(b2t_solver* s);
int b2t_stage_kkt(b2t_solver* s, double rho, int method);
int b2t_stage_pcg(b2t_solver* s, int method, double tol, int max_iter, int* iters_out /* [batch] host */);
/* standalone use of the PCG kernel (PCG(A, b, block_size = nx, Nblocks = N).solve(), PCG.py:5,214): upload a block-tridiagonal
 * system, knot-major host doubles: Sd [batch][N][nx*nx] diagonal blocks, So [batch][N][nx*nx] sub-diagonal blocks S_{k,k-1}
 * (block 0 ignored), gamma [batch][N][nx]; then b2t_stage_precond + b2t_stage_pcg, result via b2t_fetch(B2T_ARR_L). */
int b2t_set_block_system(b2t_solver* s, const double* Sd, const double* So, const double* gamma);
/* preconditioner blocks of the current S for `method` (PCG.compute_preconditioner, PCG.py:113-212) */
int b2t_stage_precond(b2t_solver* s, int method);
int b2t_stage_recover(b2t_solver* s);
int b2t_stage_merit(b2t_solver* s, double alpha, double* J, double* c, double* D /* [batch] host each */);
int b2t_fetch(b2t_solver* s, int which, double* out);

/* roofline denominator: measured FMA throughput of the device's fp64 (dtype 0) or fp32 (dtype 1) pipe in TFLOP/s
 * (148 SMs x resident blocks of independent FMA chains, CUDA-event timed, best of 5) */
int b2t_measure_fma_peak(int device, int dtype, double* tflops);

#ifdef __cplusplus
}
#endif
#endif
