// b2t_lib.cu -- host side of libb2t_<robot>.so: workspace, launch sequence of one batched SQP solve, C ABI (include/b2t.h).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>
#include "b2t_kernels.cuh"
#include "b2t_pcg_tm.cuh"
#include "b2t_ilqr.cuh"
#include "../../include/b2t.h"

// The library is compiled as two translation units so that the double and the float instantiations of every kernel build in
// parallel (B2T_PART = 1: double solver + the C ABI, B2T_PART = 2: float solver; undefined: everything in one unit).
#if !defined(B2T_PART)
#define B2T_PART 0
#endif
#define B2T_HIDDEN __attribute__((visibility("hidden")))
B2T_HIDDEN std::string& b2t_err_slot();         // thread-local last-error string (defined in the unit that holds the C ABI)
#if B2T_PART != 2
std::string& b2t_err_slot() { thread_local std::string e; return e; }
#endif

namespace {
int fail(int code, const std::string& msg) { b2t_err_slot() = msg; return code; }

#define B2T_CUDA(call)                                                                              \
  do {                                                                                              \
    cudaError_t e__ = (call);                                                                       \
    if (e__ != cudaSuccess)                                                                         \
      return fail(B2T_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__));               \
  } while (0)

inline unsigned cdiv(size_t a, size_t b) { return (unsigned)((a + b - 1) / b); }
}  // namespace

struct SolverBase {
  virtual ~SolverBase() {}
  virtual int init(const b2t_problem_desc* d, int device) = 0;
  virtual int set_trajectory(const double* x, const double* u, int on_device) = 0;
  virtual int set_goals(const double* xg, int on_device) = 0;
  virtual int set_initial_state(const double* xs) = 0;
  virtual int set_multipliers(const double* mu, const double* lam, const double* phi) = 0;
  virtual int reset_multipliers() = 0;
  virtual int solve(int method, const b2t_options* o) = 0;
  virtual int solve_ilqr(const b2t_options* o) = 0;
  virtual int mpc_shift(const double* x_next, double* x0_out, double* u0_out, double* xnext_out) = 0;
  virtual int get_trajectory(double* x, double* u, int on_device) = 0;
  virtual int get_status(int* st) = 0;
  virtual int get_scalars(double* sc) = 0;
  virtual int get_trace(double* tr, int cap) = 0;
  virtual int get_multipliers(double* mu, double* lam, double* phi) = 0;
  virtual int solve_host(const double* x0, const double* u0, const double* xg, int method, const b2t_options* o, double* xo,
                         double* uo, int* st) = 0;
  virtual int stage_dynamics() = 0;
  virtual int stage_kkt(double rho, int method) = 0;
  virtual int stage_pcg(int method, double tol, int max_iter, int* iters) = 0;
  virtual int stage_recover() = 0;
  virtual int set_block_system(const double* Sd, const double* So, const double* gam) = 0;
  virtual int stage_precond(int method) = 0;
  virtual int stage_merit(double alpha, double* J, double* c, double* D) = 0;
  virtual int fetch(int which, double* out) = 0;
  virtual int get_pass_trace(int* counts, int cap, int* passes) = 0;
  virtual const char* pcg_kernel_name() = 0;
  b2t_iteration_hook hook = nullptr;
  void* hook_user = nullptr;
  size_t ws_bytes = 0;
  cudaStream_t stream = 0;
  long long launches = 0;
  double device_seconds = 0;
  bool profiling = false;
  unsigned profile_mask = 0;       // bit f set: launches of kernel family f are bracketed by CUDA events
  double fam_seconds[B2T_KERNEL_FAMILIES] = {0};
  long long fam_launches[B2T_KERNEL_FAMILIES] = {0};
};
B2T_HIDDEN SolverBase* b2t_make_solver_f64();
B2T_HIDDEN SolverBase* b2t_make_solver_f32();
B2T_HIDDEN int b2t_fma_peak_f64(int device, double* tflops);
B2T_HIDDEN int b2t_fma_peak_f32(int device, double* tflops);

namespace {

template <typename T>
struct SolverT : SolverBase {
  b2t::Dev<T> d;
  int device = 0;
  std::vector<void*> allocs;
  double* stage_x = nullptr; double* stage_u = nullptr; double* stage_g = nullptr; double* stage_out = nullptr;
  size_t stage_out_bytes = 0;
  T* ilqr_scratch = nullptr; T* ilqr_scratch_owned = nullptr;      // trial trajectories of the parallel iLQR line search
  int* h_count = nullptr;        // pinned
  enum { COUNT_RING = 4 };
  cudaEvent_t ev_count[COUNT_RING] = {nullptr, nullptr, nullptr, nullptr};
  bool schur_v1 = false;
  int schur_minb = 2;
  int* d_status = nullptr; double* d_scalars = nullptr;
  T* d_trial_out = nullptr; int* d_trial_done = nullptr;      // k_linesearch_par: (J, c, D) per (instance, trial), arrival counters
  int ls_par_max = 160;            // passes with at most this many active instances evaluate all trials of a search at once (B2T_LS_PAR)
  bool gh_dense = false;           // d.Gh holds m x m blocks (dense KKT path / iLQR) instead of the structured path's 2m+1 scalars
  int ensure_dense_gh() {
    if (gh_dense) return 0;
    T* q = nullptr;
    if (int r = alloc(&q, (size_t)b2t::NM * b2t::NM * d.K)) return r;
    d.Gh = q;                      // the small array stays in `allocs` and is freed with the handle
    gh_dense = true;
    return 0;
  }
  int* d_scratch = nullptr;
  int* d_ticket = nullptr;         // k_pcg_tm: work-queue ticket (next slot of the work list), zeroed before every launch
  int sm_count = 148;
  bool tm_generic = false;
  bool tm_pre = true;              // A/B switch (B2T_PCG_TM_PRE=0): no TMEM prefetch ahead of the barriers
  int tm_min = 149;                // k_pcg_tm (two instances per SM) only pays when more instances are active than there are SMs (B2T_PCG_TM_MIN)
  enum { PASS_TRACE_CAP = 2048 };
  int* d_pass_trace = nullptr;      // active-instance count after every pass of the last solve (written by k_compact)
  int n_passes = 0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  std::vector<cudaEvent_t> ev_pool; size_t ev_used = 0;
  std::vector<int> ev_family;

  ~SolverT() override {
    cudaSetDevice(device);
    for (void* p : allocs) cudaFree(p);
    if (h_count) cudaFreeHost(h_count);
    if (ev0) cudaEventDestroy(ev0);
    if (ev1) cudaEventDestroy(ev1);
    for (auto e : ev_count) if (e) cudaEventDestroy(e);
    for (auto e : ev_pool) cudaEventDestroy(e);
  }

  template <typename U>
  int alloc(U** p, size_t n) {
    void* q = nullptr;
    size_t bytes = std::max<size_t>(n, 1) * sizeof(U);
    cudaError_t e = cudaMalloc(&q, bytes);
    if (e != cudaSuccess) return fail(B2T_ERR_NOMEM, std::string("cudaMalloc: ") + cudaGetErrorString(e));
    cudaMemsetAsync(q, 0, bytes, stream);
    allocs.push_back(q);
    ws_bytes += bytes;
    *p = (U*)q;
    return 0;
  }
#define B2T_ALLOC(p, n) do { int r__ = alloc(&(p), (n)); if (r__) return r__; } while (0)

  template <typename U>
  int upload(U** dst, const double* src, size_t n) {
    std::vector<U> h(n);
    for (size_t i = 0; i < n; ++i) h[i] = (U)src[i];
    B2T_ALLOC(*dst, n);
    B2T_CUDA(cudaMemcpy(*dst, h.data(), n * sizeof(U), cudaMemcpyHostToDevice));
    return 0;
  }

  int init(const b2t_problem_desc* p, int dev) override {
    using namespace b2t;
    if (!p || p->batch < 1 || p->knots < 2) return fail(B2T_ERR_INVALID, "batch >= 1 and knots >= 2 required");
    if (p->integrator_type < 0 || p->integrator_type > 3)
      return fail(B2T_ERR_UNSUPPORTED, "integrator types 0 (euler), 1 (semi-implicit euler), 2 (midpoint) and 3 (rk3); 4 (rk4) raises in the reference "
                                       "(TrajoptPlant.py:259) and -1 needs a user-coded plant");
    if (p->cost_kind == B2T_COST_URDF_EE && NJ < 2)
      return fail(B2T_ERR_UNSUPPORTED, "the end-effector cost needs a planar chain of at least 2 joints");
    if (p->cost_kind != B2T_COST_QUADRATIC && p->cost_kind != B2T_COST_URDF_EE) return fail(B2T_ERR_INVALID, "cost_kind");
    if (!p->Q || !p->QF || !p->R) return fail(B2T_ERR_INVALID, "Q, QF, R required");
    if ((size_t)p->knots * NX > (size_t)PCG_MAX_RPT * 1024) return fail(B2T_ERR_UNSUPPORTED, "knots * nx too large for the PCG block");
    device = dev;
    B2T_CUDA(cudaSetDevice(device));
    memset(&d, 0, sizeof(d));
    d.B = p->batch; d.N = p->knots; d.integrator = p->integrator_type;
    d.dt = (T)p->dt; d.gravity = (T)p->gravity;
    d.K = (size_t)d.B * d.N;
    const size_t K = d.K, B = d.B, R = (size_t)d.N * NX;
    B2T_ALLOC(d.x, NX * K); B2T_ALLOC(d.u, NU * K); B2T_ALLOC(d.xn, NX * K); B2T_ALLOC(d.un, NU * K);
    B2T_ALLOC(d.xkp1, NX * K); B2T_ALLOC(d.xkp1n, NX * K); B2T_ALLOC(d.dyn, (size_t)NDYN * K); B2T_ALLOC(d.vaf, (size_t)NVAF * K);
    B2T_ALLOC(d.g, NM * K); B2T_ALLOC(d.Gg, NM * K); B2T_ALLOC(d.dz, NM * K);      // d.Gh: after the structured / dense decision below
    B2T_ALLOC(d.Sd, B * R * NX); B2T_ALLOC(d.So, B * R * NX); B2T_ALLOC(d.Pd, B * R * NX); B2T_ALLOC(d.gam, B * R); B2T_ALLOC(d.l, B * R);
    B2T_ALLOC(d.xs, NX * B); B2T_ALLOC(d.xg, NX * B);
    // cost
    d.cost.kind = p->cost_kind; d.cost.qf_start = p->qf_start;
    if (p->hess_mode != 0 && p->hess_mode != 1)
      return fail(B2T_ERR_UNSUPPORTED, "hess_mode must be 0 (Gauss-Newton) or 1 (exact); modes 2 and 3 crash in the reference (TrajoptCost.py:500-503)");
    d.cost.hess_mode = p->cost_kind == B2T_COST_URDF_EE ? p->hess_mode : 0;
    {   // structured fast path when the quadratic cost is diagonal (B2T_DENSE_KKT=1 forces the general kernels)
      bool diag = p->cost_kind == B2T_COST_QUADRATIC;
      for (int i = 0; i < NX && diag; ++i)
        for (int j = 0; j < NX; ++j)
          if (i != j && (p->Q[i * NX + j] != 0.0 || p->QF[i * NX + j] != 0.0)) { diag = false; break; }
      for (int i = 0; i < NU && diag; ++i)
        for (int j = 0; j < NU; ++j)
          if (i != j && p->R[i * NU + j] != 0.0) { diag = false; break; }
      const char* e = getenv("B2T_DENSE_KKT");
      if (e && atoi(e) != 0) diag = false;
      d.diag_mode = diag ? 1 : 0;
    }
    { T* q; int r; if ((r = upload(&q, p->Q, NX * NX))) return r; d.cost.Q = q; }
    { T* q; int r; if ((r = upload(&q, p->QF, NX * NX))) return r; d.cost.QF = q; }
    { T* q; int r; if ((r = upload(&q, p->R, NU * NU))) return r; d.cost.R = q; }
    // limits
    int mode[NM], hmode[NM]; double lb[NM], ub[NM];
    int any = 0, any_hard = 0;
    for (int i = 0; i < NM; ++i) {
      const int ty = i < NJ ? 0 : (i < NX ? 1 : 2);
      mode[i] = p->limit_mode[ty];
      hmode[i] = 0;
      if (mode[i] < 0 || mode[i] > 3) return fail(B2T_ERR_INVALID, "limit_mode");
      if (mode[i] != B2T_LIMIT_NONE) {
        if (!p->lower || !p->upper) return fail(B2T_ERR_INVALID, "lower/upper bounds required");
        lb[i] = p->lower[i]; ub[i] = p->upper[i];
        if (mode[i] == B2T_LIMIT_ACTIVE_SET) { hmode[i] = LIM_ACTIVE_SET; mode[i] = B2T_LIMIT_NONE; any_hard = 1; }
        else any = 1;
      } else { lb[i] = 0; ub[i] = 0; }
    }
    d.lim.any = any;
    d.hard.any = any_hard;
    if (any_hard) d.diag_mode = 0;        // the elimination of the fixed coordinates lives in the dense KKT kernel
    if (d.integrator >= 2) {              // [A B] is a product of stage matrices: general kernels, [A B] stored in full
      d.diag_mode = 0;
      B2T_ALLOC(d.ABf, (size_t)NX * NM * K);
    }
    // Ghat_k: 2m+1 scalars per knot on the structured path (dinv, h, s), a dense m x m block otherwise.  iLQR always needs the dense
    // block and allocates it on its first call (ensure_dense_gh).
    gh_dense = d.diag_mode == 0;
    B2T_ALLOC(d.Gh, (size_t)(gh_dense ? NM * NM : 2 * NM + 1) * K);
    { int* m; B2T_ALLOC(m, NM); B2T_CUDA(cudaMemcpy(m, hmode, sizeof(hmode), cudaMemcpyHostToDevice)); d.hard.mode = m; }
    { int* m; B2T_ALLOC(m, NM); B2T_CUDA(cudaMemcpy(m, mode, sizeof(mode), cudaMemcpyHostToDevice)); d.lim.mode = m; }
    { T* q; int r; if ((r = upload(&q, lb, NM))) return r; d.lim.lb = q; }
    { T* q; int r; if ((r = upload(&q, ub, NM))) return r; d.lim.ub = q; }
    d.hard.lb = d.lim.lb; d.hard.ub = d.lim.ub;
    for (int ty = 0; ty < 3; ++ty) {
      d.mu_init[ty] = (T)p->mu_init[ty]; d.mu_factor[ty] = (T)p->mu_factor[ty]; d.mu_max[ty] = (T)p->mu_max[ty];
      d.phi_init[ty] = (T)p->phi_init[ty]; d.phi_factor[ty] = (T)p->phi_factor[ty];
    }
    const size_t nmult = any ? (size_t)2 * NM * K : 1;
    B2T_ALLOC(d.mu, nmult); B2T_ALLOC(d.lam, nmult); B2T_ALLOC(d.phi, nmult);
    // per-instance state
    B2T_ALLOC(d.rho, B); B2T_ALLOC(d.drho, B); B2T_ALLOC(d.J, B); B2T_ALLOC(d.c, B); B2T_ALLOC(d.merit, B); B2T_ALLOC(d.alpha, B);
    B2T_ALLOC(d.deltaJ, B); B2T_ALLOC(d.D, B); B2T_ALLOC(d.ratio, B);
    B2T_ALLOC(d.ls_iter, B); B2T_ALLOC(d.sqp_iter, B); B2T_ALLOC(d.outer_iter, B); B2T_ALLOC(d.exit_sqp, B); B2T_ALLOC(d.exit_soft, B);
    B2T_ALLOC(d.phase, B); B2T_ALLOC(d.dyn_ok, B); B2T_ALLOC(d.err, B); B2T_ALLOC(d.pcg_iters, B); B2T_ALLOC(d.tot_qp, B); B2T_ALLOC(d.tot_pcg, B); B2T_ALLOC(d.tot_trials, B);
    B2T_ALLOC(d.act, B); B2T_ALLOC(d.n_act, 1); B2T_ALLOC(d.ls_list0, B); B2T_ALLOC(d.ls_list1, B); B2T_ALLOC(d.restart_list, B); B2T_ALLOC(d.n_restart, 1);
    B2T_ALLOC(d.n_ls, MAX_LS_TRIALS + 1);
    B2T_ALLOC(d.nu_trace, B * NU_TRACE_LEN);
    d.trace_cap = 104;
    B2T_ALLOC(d.trace, B * d.trace_cap * TRACE_FIELDS); B2T_ALLOC(d.trace_rows, B);
    B2T_ALLOC(d_scratch, B); B2T_ALLOC(d_ticket, 4); B2T_ALLOC(d_status, B * 8); B2T_ALLOC(d_scalars, B * 4); B2T_ALLOC(d_pass_trace, PASS_TRACE_CAP);
    B2T_ALLOC(stage_x, (size_t)B * NX * d.N); B2T_ALLOC(stage_u, (size_t)B * NU * (d.N - 1)); B2T_ALLOC(stage_g, (size_t)B * NX);
    stage_out_bytes = std::max<size_t>({(size_t)NM * NM * K, (size_t)NDYN * K, (size_t)2 * NM * K, B * (size_t)d.trace_cap * TRACE_FIELDS,
                                        (size_t)(2 * NX * NX + NX) * K}) * sizeof(double);
    { void* q; cudaError_t e = cudaMalloc(&q, stage_out_bytes); if (e != cudaSuccess) return fail(B2T_ERR_NOMEM, "cudaMalloc stage_out"); allocs.push_back(q); ws_bytes += stage_out_bytes; stage_out = (double*)q; }
    B2T_CUDA(cudaMallocHost((void**)&h_count, 64));
    B2T_CUDA(cudaEventCreate(&ev0)); B2T_CUDA(cudaEventCreate(&ev1));
    for (int i = 0; i < COUNT_RING; ++i) B2T_CUDA(cudaEventCreateWithFlags(&ev_count[i], cudaEventDisableTiming));
    // kernels that need > 48 KB of dynamic shared memory
    B2T_CUDA(cudaFuncSetAttribute(b2t::k_pcg<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    B2T_CUDA(cudaFuncSetAttribute(b2t::k_linesearch<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    B2T_CUDA(cudaFuncSetAttribute(b2t::k_linesearch_par<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    B2T_CUDA(cudaFuncSetAttribute(b2t::k_linesearch<T, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    B2T_CUDA(cudaFuncSetAttribute(b2t::k_linesearch_par<T, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    B2T_ALLOC(d_trial_out, (size_t)B * MAX_LS_TRIALS * 3); B2T_ALLOC(d_trial_done, B);
    { const char* e = getenv("B2T_LS_PAR"); if (e) ls_par_max = atoi(e); }
    B2T_CUDA(cudaFuncSetAttribute(b2t::k_schur_diag<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)NJ * NM * SCHUR_THREADS * sizeof(T))));
    B2T_CUDA(cudaFuncSetAttribute(b2t::k_schur_rows<T, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)SCHUR_REC * SCHUR_KB * sizeof(T))));
    B2T_CUDA(cudaFuncSetAttribute(b2t::k_schur_rows<T, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)SCHUR_REC * SCHUR_KB * sizeof(T))));
    { const char* e = getenv("B2T_SCHUR_MINB"); schur_minb = (e && atoi(e) == 3) ? 3 : 2; }
    { const char* e = getenv("B2T_SCHUR_V1"); schur_v1 = e && atoi(e) != 0; }     // A/B switch: the one-thread-per-block-row kernel
    B2T_CUDA(cudaFuncSetAttribute(b2t::k_pcg2<T, PCG_RPT, 1, true, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    B2T_CUDA(cudaFuncSetAttribute(b2t::k_pcg2<T, PCG_RPT, 1, true, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    B2T_CUDA(cudaFuncSetAttribute(b2t::k_pcg2<T, PCG_RPT, 1, true, 1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    // the streaming variant (SM = false) keeps 3 (N+2) nx scalars of vectors in shared memory: beyond 48 KB for long horizons
    B2T_CUDA(cudaFuncSetAttribute(b2t::k_pcg2<T, PCG_RPT, 1, false, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    B2T_CUDA(cudaFuncSetAttribute(b2t::k_pcg2<T, PCG_RPT, 1, false, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    B2T_CUDA(cudaFuncSetAttribute(b2t::k_pcg2<T, PCG_RPT, 1, false, 1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    if constexpr (b2t::NX % 4 == 0) {
      B2T_CUDA(cudaFuncSetAttribute(b2t::k_pcg3<T, 256, false, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
      B2T_CUDA(cudaFuncSetAttribute(b2t::k_pcg3<T, 512, false, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
      B2T_CUDA(cudaFuncSetAttribute(b2t::k_pcg3<T, 1024, false, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
      B2T_CUDA(cudaFuncSetAttribute(b2t::k_pcg3<T, 128, true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024));
      if constexpr (sizeof(T) == 8 && b2t::pcg_tm_eligible())
      {
        B2T_CUDA(cudaFuncSetAttribute(b2t::k_pcg_tm<T, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
        B2T_CUDA(cudaFuncSetAttribute(b2t::k_pcg_tm<T, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
        B2T_CUDA(cudaFuncSetAttribute(b2t::k_pcg_tm<T, 256, 64, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
        B2T_CUDA(cudaFuncSetAttribute(b2t::k_pcg_tm<T, 256, 64, 0, 2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
      }
    }
    B2T_CUDA(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, device));
    tm_min = sm_count + 1;
    { const char* e = getenv("B2T_PCG_TM_MIN"); if (e) tm_min = atoi(e); }
    { const char* e = getenv("B2T_PCG_TM_PRE"); if (e) tm_pre = atoi(e) != 0; }
    { const char* e = getenv("B2T_PCG_TM_GENERIC"); tm_generic = e && atoi(e) != 0; }     // A/B switch: run-time horizon / integrator instantiation
    if constexpr (PCG_CS_MAX == 2) {
      B2T_CUDA(cudaFuncSetAttribute(b2t::k_pcg2<T, PCG_RPT, 2, true, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      B2T_CUDA(cudaFuncSetAttribute(b2t::k_pcg2<T, PCG_RPT, 2, true, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      B2T_CUDA(cudaFuncSetAttribute(b2t::k_pcg2<T, PCG_RPT, 2, true, 1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      B2T_CUDA(cudaFuncSetAttribute(b2t::k_pcg2<T, PCG_RPT, 2, false, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      B2T_CUDA(cudaFuncSetAttribute(b2t::k_pcg2<T, PCG_RPT, 2, false, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      B2T_CUDA(cudaFuncSetAttribute(b2t::k_pcg2<T, PCG_RPT, 2, false, 1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    }
    int r = reset_multipliers();
    if (r) return r;
    k_init_state<T><<<cdiv(B, 128), 128, 0, stream>>>(d);
    B2T_CUDA(cudaGetLastError());
    B2T_CUDA(cudaStreamSynchronize(stream));
    return 0;
  }

  // ------------------------------------------------------------------ launch bookkeeping
  void tick(int family) {
    ++launches;
    ++fam_launches[family];
  }
  struct Scope {
    SolverT* s; int fam; cudaEvent_t a = nullptr, b = nullptr;
    Scope(SolverT* s_, int fam_) : s(s_), fam(fam_) {
      if ((s->profile_mask >> fam) & 1u) { a = s->next_event(); b = s->next_event(); s->ev_family.push_back(fam); cudaEventRecord(a, s->stream); }
    }
    ~Scope() { if (b) cudaEventRecord(b, s->stream); }
  };
  cudaEvent_t next_event() {
    if (ev_used == ev_pool.size()) { cudaEvent_t e; cudaEventCreate(&e); ev_pool.push_back(e); }
    return ev_pool[ev_used++];
  }
  void collect_profile() {
    if (!profiling) return;
    cudaStreamSynchronize(stream);
    for (size_t i = 0; i < ev_family.size(); ++i) {
      float ms = 0;
      cudaEventElapsedTime(&ms, ev_pool[2 * i], ev_pool[2 * i + 1]);
      fam_seconds[ev_family[i]] += ms * 1e-3;
    }
    ev_used = 0; ev_family.clear();
  }

  // ------------------------------------------------------------------ inputs / outputs
  int set_trajectory(const double* x, const double* u, int on_device) override {
    using namespace b2t;
    if (!x || !u) return fail(B2T_ERR_INVALID, "x, u required");
    B2T_CUDA(cudaSetDevice(device));
    const double *dx = x, *du = u;
    if (!on_device) {
      B2T_CUDA(cudaMemcpyAsync(stage_x, x, (size_t)d.B * NX * d.N * sizeof(double), cudaMemcpyHostToDevice, stream));
      B2T_CUDA(cudaMemcpyAsync(stage_u, u, (size_t)d.B * NU * (d.N - 1) * sizeof(double), cudaMemcpyHostToDevice, stream));
      dx = stage_x; du = stage_u;
    }
    k_pack_traj<T><<<cdiv(d.K, 128), 128, 0, stream>>>(d, dx, du);
    B2T_CUDA(cudaGetLastError());
    return 0;
  }
  int set_goals(const double* xg, int on_device) override {
    using namespace b2t;
    if (!xg) return fail(B2T_ERR_INVALID, "xg required");
    B2T_CUDA(cudaSetDevice(device));
    const double* dg = xg;
    if (!on_device) {
      B2T_CUDA(cudaMemcpyAsync(stage_g, xg, (size_t)d.B * NX * sizeof(double), cudaMemcpyHostToDevice, stream));
      dg = stage_g;
    }
    k_pack_goals<T><<<cdiv(d.B, 128), 128, 0, stream>>>(d, dg);
    B2T_CUDA(cudaGetLastError());
    return 0;
  }
  int set_initial_state(const double* xs) override {
    using namespace b2t;
    if (!xs) return fail(B2T_ERR_INVALID, "xs required");
    B2T_CUDA(cudaSetDevice(device));
    std::vector<T> h((size_t)NX * d.B);
    for (int b = 0; b < d.B; ++b)
      for (int i = 0; i < NX; ++i) h[(size_t)i * d.B + b] = (T)xs[(size_t)b * NX + i];
    B2T_CUDA(cudaStreamSynchronize(stream));
    B2T_CUDA(cudaMemcpy(d.xs, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
    return 0;
  }
  int set_multipliers(const double* mu, const double* lam, const double* phi) override {
    using namespace b2t;
    if (!d.lim.any) return fail(B2T_ERR_INVALID, "no soft limits configured");
    B2T_CUDA(cudaSetDevice(device));
    const size_t bytes = (size_t)2 * NM * d.K * sizeof(double);
    const double* src[3] = {mu, lam, phi};
    T* dst[3] = {d.mu, d.lam, d.phi};
    for (int i = 0; i < 3; ++i) {
      if (!src[i]) continue;
      B2T_CUDA(cudaMemcpyAsync(stage_out, src[i], bytes, cudaMemcpyHostToDevice, stream));
      k_pack_mult<T><<<cdiv(d.K, 128), 128, 0, stream>>>(d, stage_out, dst[i]);
      B2T_CUDA(cudaGetLastError());
      B2T_CUDA(cudaStreamSynchronize(stream));
    }
    return 0;
  }
  int get_multipliers(double* mu, double* lam, double* phi) override {
    using namespace b2t;
    if (!d.lim.any) return fail(B2T_ERR_INVALID, "no soft limits configured");
    B2T_CUDA(cudaSetDevice(device));
    const size_t bytes = (size_t)2 * NM * d.K * sizeof(double);
    double* dst[3] = {mu, lam, phi};
    T* src[3] = {d.mu, d.lam, d.phi};
    for (int i = 0; i < 3; ++i) {
      if (!dst[i]) continue;
      k_unpack_mult<T><<<cdiv(d.K, 128), 128, 0, stream>>>(d, src[i], stage_out);
      B2T_CUDA(cudaGetLastError());
      B2T_CUDA(cudaMemcpyAsync(dst[i], stage_out, bytes, cudaMemcpyDeviceToHost, stream));
      B2T_CUDA(cudaStreamSynchronize(stream));
    }
    return 0;
  }
  int reset_multipliers() override {
    using namespace b2t;
    if (!d.lim.any) return 0;
    B2T_CUDA(cudaSetDevice(device));
    k_init_mult<T><<<cdiv(d.K, 128), 128, 0, stream>>>(d);
    B2T_CUDA(cudaGetLastError());
    return 0;
  }
  int get_trajectory(double* x, double* u, int on_device) override {
    using namespace b2t;
    if (!x || !u) return fail(B2T_ERR_INVALID, "x, u required");
    B2T_CUDA(cudaSetDevice(device));
    double* dx = on_device ? x : stage_x;
    double* du = on_device ? u : stage_u;
    k_unpack_traj<T><<<cdiv(d.K, 128), 128, 0, stream>>>(d, dx, du);
    B2T_CUDA(cudaGetLastError());
    if (!on_device) {
      B2T_CUDA(cudaMemcpyAsync(x, stage_x, (size_t)d.B * NX * d.N * sizeof(double), cudaMemcpyDeviceToHost, stream));
      B2T_CUDA(cudaMemcpyAsync(u, stage_u, (size_t)d.B * NU * (d.N - 1) * sizeof(double), cudaMemcpyDeviceToHost, stream));
    }
    B2T_CUDA(cudaStreamSynchronize(stream));
    return 0;
  }
  int pack_status() {
    using namespace b2t;
    k_pack_status<T><<<cdiv(d.B, 128), 128, 0, stream>>>(d, d_status, d_scalars);
    B2T_CUDA(cudaGetLastError());
    return 0;
  }
  int get_status(int* st) override {
    if (!st) return fail(B2T_ERR_INVALID, "status required");
    B2T_CUDA(cudaSetDevice(device));
    int r = pack_status(); if (r) return r;
    B2T_CUDA(cudaMemcpyAsync(st, d_status, (size_t)d.B * 8 * sizeof(int), cudaMemcpyDeviceToHost, stream));
    B2T_CUDA(cudaStreamSynchronize(stream));
    return 0;
  }
  int get_scalars(double* sc) override {
    if (!sc) return fail(B2T_ERR_INVALID, "scalars required");
    B2T_CUDA(cudaSetDevice(device));
    int r = pack_status(); if (r) return r;
    B2T_CUDA(cudaMemcpyAsync(sc, d_scalars, (size_t)d.B * 4 * sizeof(double), cudaMemcpyDeviceToHost, stream));
    B2T_CUDA(cudaStreamSynchronize(stream));
    return 0;
  }
  int get_trace(double* tr, int cap) override {
    using namespace b2t;
    if (!tr || cap < 1) return fail(B2T_ERR_INVALID, "trace buffer required");
    B2T_CUDA(cudaSetDevice(device));
    const size_t n = (size_t)d.B * cap * TRACE_FIELDS;
    if (n * sizeof(double) > stage_out_bytes) return fail(B2T_ERR_INVALID, "trace_cap too large");
    k_fetch_trace<T><<<cdiv(n, 256), 256, 0, stream>>>(d, stage_out, cap);
    B2T_CUDA(cudaGetLastError());
    B2T_CUDA(cudaMemcpyAsync(tr, stage_out, n * sizeof(double), cudaMemcpyDeviceToHost, stream));
    B2T_CUDA(cudaStreamSynchronize(stream));
    return 0;
  }

  // ------------------------------------------------------------------ kernels of one QP solve
  static b2t::Opts<T> convert(const b2t_options* o) {
    b2t::Opts<T> r;
    r.tol_lin = (T)o->exit_tolerance_linSys; r.max_iter_lin = o->max_iter_linSys;
    r.tol_sqp = (T)o->exit_tolerance_SQP; r.max_iter_sqp = o->max_iter_SQP;
    r.alpha_factor = (T)o->alpha_factor; r.alpha_min = (T)o->alpha_min;
    r.rho_factor = (T)o->rho_factor; r.rho_min = (T)o->rho_min; r.rho_max = (T)o->rho_max; r.rho_init = (T)o->rho_init;
    r.er_min = (T)o->expected_reduction_min; r.er_max = (T)o->expected_reduction_max;
    r.tol_soft = (T)o->exit_tolerance_soft; r.max_iter_soft = o->max_iter_soft;
    r.merit_mu = (T)o->merit_mu;
    return r;
  }
  size_t pcg_smem() const { return ((size_t)3 * d.N * b2t::NX + 64) * sizeof(T); }
  int pcg_threads() const {
    int R = d.N * b2t::NX;
    int t = ((R + 31) / 32) * 32;
    return std::min(t, 1024);
  }
  int merit_threads() const { return std::min(256, ((d.N + 31) / 32) * 32); }

  int launch_dynamics(const int* list, const int* count, int bound) {
    using namespace b2t;
    const size_t nthreads = (size_t)bound * d.N;
    if (d.integrator >= 2) { Scope sc(this, B2T_K_FD); k_fd<T, false, true><<<cdiv(nthreads, 128), 128, 0, stream>>>(d, list, count); tick(B2T_K_FD); }
    else { Scope sc(this, B2T_K_FD); k_fd<T, false><<<cdiv(nthreads, 128), 128, 0, stream>>>(d, list, count); tick(B2T_K_FD); }
    if (d.integrator >= 2) { Scope sc(this, B2T_K_FDGRAD); k_ab_multi<T><<<cdiv(nthreads, 64), 64, 0, stream>>>(d, list, count); tick(B2T_K_FDGRAD); }
    else { Scope sc(this, B2T_K_FDGRAD); k_fd_grad<T><<<(unsigned)(cdiv(nthreads, 128) * 2 * NJ), 128, 0, stream>>>(d, list, count); tick(B2T_K_FDGRAD); }
    return 0;
  }
  int launch_kkt(const int* list, const int* count, int bound, int method, bool all_outputs = false) {
    using namespace b2t;
    const size_t nthreads = (size_t)bound * d.N;
    const int jac = method == B2T_METHOD_PCG_J ? 1 : 0;
    if (d.diag_mode) {
      // the matrix-free PCG kernels never read the sub-diagonal blocks S_{k,k-1}: skip their stores (half of this kernel's traffic)
      decide_pcg_variant();
      const bool exact = method == B2T_METHOD_N || method == B2T_METHOD_S;
      const int need_so = (all_outputs || exact || pcg_variant < 3) ? 1 : 0;
      { Scope sc(this, B2T_K_KKT); k_kkt_diag<T><<<cdiv(nthreads, 128), 128, 0, stream>>>(d, list, count); tick(B2T_K_KKT); }
      if (schur_v1) { Scope sc(this, B2T_K_SCHUR); k_schur_diag<T><<<cdiv(nthreads, SCHUR_THREADS), SCHUR_THREADS, (size_t)NJ * NM * SCHUR_THREADS * sizeof(T), stream>>>(d, list, count, need_so); tick(B2T_K_SCHUR); }
      else {
        Scope sc(this, B2T_K_SCHUR);
        const size_t sm = (size_t)SCHUR_REC * SCHUR_KB * sizeof(T);
        if (schur_minb == 3) k_schur_rows<T, 3><<<cdiv(nthreads, SCHUR_KB), SCHUR_KB * NJ, sm, stream>>>(d, list, count, need_so);
        else k_schur_rows<T, 2><<<cdiv(nthreads, SCHUR_KB), SCHUR_KB * NJ, sm, stream>>>(d, list, count, need_so);
        tick(B2T_K_SCHUR);
      }
      if (method != B2T_METHOD_N && method != B2T_METHOD_S) {
        Scope sc(this, B2T_K_SCHUR); k_pinv<T><<<cdiv(nthreads, 128), 128, 0, stream>>>(d, list, count, jac); tick(B2T_K_SCHUR);
      }
      return 0;
    }
    { Scope sc(this, B2T_K_KKT); k_kkt<T><<<cdiv(nthreads, 64), 64, 0, stream>>>(d, list, count); tick(B2T_K_KKT); }
    { Scope sc(this, B2T_K_SCHUR); k_schur<T><<<cdiv(nthreads, 64), 64, 0, stream>>>(d, list, count, jac); tick(B2T_K_SCHUR); }
    return 0;
  }
  // rows per thread of k_pcg2: largest divisor of NX that is <= 3
  static constexpr bool pcg_rpt_ok(int r) { int tb = b2t::NX / r; return b2t::NX % r == 0 && (tb & (tb - 1)) == 0 && tb <= 32; }
  static constexpr int pick_rpt() {
    if (pcg_rpt_ok(3)) return 3;
    for (int r = 2; r <= b2t::NX; ++r) if (pcg_rpt_ok(r)) return r;
    return b2t::NX;
  }
  static constexpr int PCG_RPT = pick_rpt();
  // column split: 2 when every thread still gets an even number of columns and the block fits 1024 threads
  static constexpr int PCG_CS_MAX = (b2t::NX % 4 == 0 && 32 % ((b2t::NX / PCG_RPT) * 2) == 0) ? 2 : 1;
  int pcg_cs = 0;
  int pcg2_cs() {
    if (!pcg_cs) {
      const char* e = getenv("B2T_PCG_CS");
      int want = e ? atoi(e) : 1;      // measured on B200 (arm6, N=64): CS=1 22.3 ms vs CS=2 28.3 ms per 2048-instance step
      if (want > PCG_CS_MAX) want = PCG_CS_MAX;
      if (want == 2 && d.N * (b2t::NX / PCG_RPT) * 2 > 1024) want = 1;
      pcg_cs = want < 1 ? 1 : want;
    }
    return pcg_cs;
  }
  int pcg2_threads() { int nt = d.N * (b2t::NX / PCG_RPT) * pcg2_cs(); return ((nt + 31) / 32) * 32; }
  int pcg2_maxt() { int nt = pcg2_threads(); return nt <= 256 ? 256 : (nt <= 512 ? 512 : 1024); }
  size_t pcg2_smem(bool mats) {
    size_t v = ((size_t)3 * (d.N + 2) * b2t::NX + 32) * sizeof(T);
    if (mats) v += (size_t)2 * PCG_RPT * (b2t::NX / pcg2_cs()) * pcg2_maxt() * sizeof(T);
    return v;
  }
  // 0: k_pcg (v1), 1: k_pcg2 with shared-memory diagonal blocks (default), 2: k_pcg2 streaming them from L1/L2,
  // 3: k_pcg3 matrix-free, register-resident (default for the structured path when 4 N <= 256), 4: k_pcg3 with two lanes per knot
  // and the preconditioner rows in shared memory (two instances per SM; measured 24.4 ms)
  int pcg_variant = -1;
  bool pcg_col = false;
  bool explicit_system = false;   // set by b2t_set_block_system: only S / Pinv blocks are valid -> the explicit kernels must run
  // k_pcg_tm: fp64, structured path, one instance per 256 threads, D^-1 rows within the per-thread TMEM window
  bool tm_ok() const { return sizeof(T) == 8 && b2t::pcg_tm_eligible() && d.diag_mode && 4 * d.N <= b2t::PCGTM_THREADS; }
  void decide_pcg_variant() {
    if (pcg_variant < 0) {
      const char* e = getenv("B2T_PCG_VARIANT");
      if (e) {
        pcg_variant = atoi(e);
        if (pcg_variant == 7) { pcg_variant = 3; pcg_col = true; }      // experiment: k_pcg3 with the column form of the D^-1 products
      }
      // k_pcg3: 21.9 ms vs 24.0 ms (variant 1) per 2048-instance step at N = 64.  Longer horizons use its 512- / 1024-thread
      // instantiations (128 / 64 registers, spilling): still 1.85x (N = 128: 76.7 vs 141.8 ns per instance-iteration) and 3.2x
      // (N = 256: 289 vs 932 ns) faster than the explicit-block kernel, which no longer fits its blocks in shared memory there
      else if (tm_ok() && d.N > 32) pcg_variant = 8;      // shorter horizons: k_pcg3 launches 4 N threads and already fits 2 - 4 CTAs per SM
      else if (d.diag_mode && b2t::NX % 4 == 0 && 4 * d.N <= 1024) pcg_variant = 3;
      else if (pcg2_threads() > 1024) pcg_variant = 0;
      else pcg_variant = pcg2_smem(true) <= (size_t)220 * 1024 ? 1 : 2;
      if (pcg_variant == 8 && !tm_ok()) pcg_variant = 3;
      if (pcg_variant == 6 && !(d.diag_mode && b2t::NJ % 6 == 0 && d.N <= b2t::PCG6_KPW * (b2t::PCG6_THREADS / 32))) pcg_variant = 3;
      if (pcg_variant == 5 && !(d.diag_mode && b2t::NX % 4 == 0 && 4 * d.N <= 256)) pcg_variant = 3;
      if (pcg_variant == 3 && !(d.diag_mode && b2t::NX % 4 == 0 && 4 * d.N <= 1024)) pcg_variant = 1;
      if (pcg_variant == 4 && !(d.diag_mode && b2t::NX % 4 == 0 && 2 * d.N <= 128)) pcg_variant = 1;
    }
  }
  int launch_pcg(const int* list, const int* count, int bound, int method, T tol, int max_iter) {
    using namespace b2t;
    decide_pcg_variant();
    const int saved_variant = pcg_variant;
    if (explicit_system && pcg_variant >= 3)
      pcg_variant = pcg2_threads() > 1024 ? 0 : (pcg2_smem(true) <= (size_t)220 * 1024 ? 1 : 2);
    struct Restore { int& v; int s; ~Restore() { v = s; } } restore{pcg_variant, saved_variant};
    const int stair = method == B2T_METHOD_PCG_SS ? 1 : 0;
    Scope sc(this, B2T_K_PCG);
    if (method == B2T_METHOD_N || method == B2T_METHOD_S) {
      k_bt_solve<T><<<cdiv(bound, 32), 32, 0, stream>>>(d, list, count);
      tick(B2T_K_PCG);
      return 0;
    }
    const int nt = pcg2_threads();
    if (pcg_variant == 6) {
      // six lanes per knot, five knots per warp (robots with nj % 6 == 0)
      if constexpr (b2t::NJ % 6 == 0) {
        const int nt6 = ((d.N + PCG6_KPW - 1) / PCG6_KPW) * 32;
        const size_t sm6 = ((size_t)2 * (d.N + 1) * PCG6_VS + (size_t)2 * d.N * PCG6_VS + 32) * sizeof(T);
        if (d.integrator == 0) k_pcg6<T, true><<<bound, nt6, sm6, stream>>>(d, list, count, stair, tol, max_iter);
        else k_pcg6<T, false><<<bound, nt6, sm6, stream>>>(d, list, count, stair, tol, max_iter);
      }
      tick(B2T_K_PCG);
      return 0;
    }
    // N <= 64: two instances per CTA, worth it when more instances are active than there are SMs (below that k_pcg3 has the lower
    // latency).  64 < N <= 128: one instance per CTA; k_pcg3's 512-thread instantiation spills at 128 registers, this one does not.
    if (pcg_variant == 8 && (4 * d.N > 256 || bound >= tm_min)) {
      // matrices in tensor memory, persistent CTAs that draw instances from a ticket counter
      if constexpr (sizeof(T) == 8 && b2t::pcg_tm_eligible()) {
        const size_t smh = ((size_t)2 * (d.N + 1) * NX + (size_t)2 * d.N * PCG3_NMS + 64) * sizeof(T);
        B2T_CUDA(cudaMemsetAsync(d_ticket, 0, sizeof(int), stream));
        if (d.N == 64 && d.integrator == 0 && !tm_generic && !tm_pre) k_pcg_tm<T, 256, 64, 0, 2, false><<<std::min(sm_count, (bound + 1) / 2), PCGTM_THREADS, 2 * smh, stream>>>(d, list, count, d_ticket, stair, tol, max_iter);
        else if (d.N == 64 && d.integrator == 0 && !tm_generic) k_pcg_tm<T, 256, 64, 0><<<std::min(sm_count, (bound + 1) / 2), PCGTM_THREADS, 2 * smh, stream>>>(d, list, count, d_ticket, stair, tol, max_iter);
        else if (4 * d.N <= 256) k_pcg_tm<T, 256><<<std::min(sm_count, (bound + 1) / 2), PCGTM_THREADS, 2 * smh, stream>>>(d, list, count, d_ticket, stair, tol, max_iter);
        else k_pcg_tm<T, 512><<<std::min(sm_count, bound), PCGTM_THREADS, smh, stream>>>(d, list, count, d_ticket, stair, tol, max_iter);
      }
      tick(B2T_K_PCG);
      return 0;
    }
    if (pcg_variant == 3 || pcg_variant == 4 || pcg_variant == 5 || pcg_variant == 8) {
      // 3: four lanes per knot, everything in registers, one instance per SM;  4: two lanes per knot, preconditioner rows in
      // shared memory, two instances per SM;  5: k_pcg4 = 3 with the own-block halves of the products ahead of the barriers
      if constexpr (b2t::NX % 4 == 0) {
        const size_t smv = ((size_t)2 * (d.N + 1) * NX + (size_t)2 * d.N * PCG3_NMS + 64) * sizeof(T);
        if (pcg_variant == 5) {
          const int nt3 = ((4 * d.N + 31) / 32) * 32;
          if (d.integrator == 0) k_pcg4<T, 256, true><<<bound, nt3, smv, stream>>>(d, list, count, stair, tol, max_iter);
          else k_pcg4<T, 256, false><<<bound, nt3, smv, stream>>>(d, list, count, stair, tol, max_iter);
        } else if (pcg_variant == 4) {
          const int nt4 = ((2 * d.N + 31) / 32) * 32;
          const size_t smp = smv + (size_t)(NX / 2) * NX * 128 * sizeof(T);
          k_pcg3<T, 128, true, 2><<<bound, nt4, smp, stream>>>(d, list, count, stair, tol, max_iter);
        } else {
          const int nt3 = ((4 * d.N + 31) / 32) * 32;
          if (nt3 <= 256 && pcg_col) k_pcg3<T, 256, false, 4, true><<<bound, nt3, smv, stream>>>(d, list, count, stair, tol, max_iter);
          else if (nt3 <= 256) k_pcg3<T, 256, false, 4><<<bound, nt3, smv, stream>>>(d, list, count, stair, tol, max_iter);
          else if (nt3 <= 512) k_pcg3<T, 512, false, 4><<<bound, nt3, smv, stream>>>(d, list, count, stair, tol, max_iter);
          else k_pcg3<T, 1024, false, 4><<<bound, nt3, smv, stream>>>(d, list, count, stair, tol, max_iter);
        }
      }
      tick(B2T_K_PCG);
      return 0;
    }
    const int cs = pcg2_cs();
#define B2T_PCG2(CSV, SM, MT) k_pcg2<T, PCG_RPT, CSV, SM, MT><<<bound, nt, pcg2_smem(SM), stream>>>(d, list, count, stair, tol, max_iter)
#define B2T_PCG2_MT(CSV, SM) do { if (nt <= 256) B2T_PCG2(CSV, SM, 256); else if (nt <= 512) B2T_PCG2(CSV, SM, 512); else B2T_PCG2(CSV, SM, 1024); } while (0)
    if (pcg_variant == 1 || pcg_variant == 2) {
      const bool sm = pcg_variant == 1;
      if (cs == 2) { if constexpr (PCG_CS_MAX == 2) { if (sm) B2T_PCG2_MT(2, true); else B2T_PCG2_MT(2, false); } }
      else { if (sm) B2T_PCG2_MT(1, true); else B2T_PCG2_MT(1, false); }
    } else k_pcg<T><<<bound, pcg_threads(), pcg_smem(), stream>>>(d, list, count, stair, tol, max_iter);
#undef B2T_PCG2_MT
#undef B2T_PCG2
    tick(B2T_K_PCG);
    return 0;
  }
  int launch_recover(const int* list, const int* count, int bound) {
    using namespace b2t;
    Scope sc(this, B2T_K_RECOVER);
    if (d.diag_mode) k_recover_diag<T><<<cdiv((size_t)bound * d.N, 128), 128, 0, stream>>>(d, list, count);
    else k_recover<T><<<cdiv((size_t)bound * d.N, 128), 128, 0, stream>>>(d, list, count);
    tick(B2T_K_RECOVER);
    return 0;
  }

  int solve(int method, const b2t_options* o) override {
    using namespace b2t;
    if (!o) return fail(B2T_ERR_INVALID, "options required");
    if (method != B2T_METHOD_N && method != B2T_METHOD_S && method != B2T_METHOD_PCG_J && method != B2T_METHOD_PCG_BJ &&
        method != B2T_METHOD_PCG_SS)
      return fail(B2T_ERR_INVALID, "method must be N, S, PCG-J, PCG-BJ or PCG-SS");
    // trace rows beyond trace_cap (104 per outer iteration) are dropped by trace_row(); iteration counts and results do not depend on them
    if (d.hard.any && method != B2T_METHOD_N && method != B2T_METHOD_S)
      return fail(B2T_ERR_UNSUPPORTED, "hard (ACTIVE_SET) limits: exact methods N / S only (the reference hands PCG a Schur complement whose size no "
                                       "longer matches block_size * Nblocks)");
    // the backtracking loop of k_linesearch ends when alpha <= alpha_min: it needs a factor that shrinks alpha
    if (!(o->alpha_factor > 0.0 && o->alpha_factor < 1.0)) return fail(B2T_ERR_INVALID, "alpha_factor_SQP_DDP must lie in (0, 1)");
    if (!(o->alpha_min > 0.0)) return fail(B2T_ERR_INVALID, "alpha_min_SQP_DDP must be positive");
    if (o->max_iter_SQP < 1 || o->max_iter_soft < 1) return fail(B2T_ERR_INVALID, "max_iter_SQP_DDP and max_iter_softConstraints must be >= 1");
    B2T_CUDA(cudaSetDevice(device));
    explicit_system = false;
    Opts<T> op = convert(o);
    int max_trials = 1;
    bool trials_exact = false;      // max_trials is the sequential loop's true bound (not the MAX_LS_TRIALS cap)
    { T a = T(1); while (a > op.alpha_min && max_trials < MAX_LS_TRIALS) { a *= op.alpha_factor; ++max_trials; } trials_exact = !(a > op.alpha_min); }
    launches = 0; device_seconds = 0;
    for (int i = 0; i < B2T_KERNEL_FAMILIES; ++i) { fam_seconds[i] = 0; fam_launches[i] = 0; }
    B2T_CUDA(cudaEventRecord(ev0, stream));
    const int B = d.B;
    const size_t msmem = (size_t)6 * d.N * sizeof(T);
    const int mt = merit_threads();
    k_init_state<T><<<cdiv(B, 128), 128, 0, stream>>>(d); tick(B2T_K_CTRL);
    if (d.integrator >= 2) { Scope sc(this, B2T_K_FD); k_fd<T, false, true><<<cdiv((size_t)B * d.N, 128), 128, 0, stream>>>(d, d.act, d.n_act); tick(B2T_K_FD); }
    else { Scope sc(this, B2T_K_FD); k_fd<T, false><<<cdiv((size_t)B * d.N, 128), 128, 0, stream>>>(d, d.act, d.n_act); tick(B2T_K_FD); }
    { Scope sc(this, B2T_K_MERIT); k_outer_begin<T><<<B, mt, msmem, stream>>>(d, d.act, d.n_act, op, 1); tick(B2T_K_MERIT); }
    B2T_CUDA(cudaGetLastError());
    int n = B;
    const long long cap = (long long)o->max_iter_soft * o->max_iter_SQP + 8;
    const bool legacy_ls = getenv("B2T_LEGACY_LS") && atoi(getenv("B2T_LEGACY_LS")) != 0;
    const bool trace_active = getenv("B2T_TRACE_ACTIVE") != nullptr;      // debugging: active-instance count after every SQP pass on stderr
    const bool lagged = !hook && !trace_active && !(getenv("B2T_SYNC_PASSES") && atoi(getenv("B2T_SYNC_PASSES")) != 0);
    const size_t lsmem = (size_t)(6 + NX) * d.N * sizeof(T);
    const size_t lsmem_par = (size_t)(6 + 2 * NX) * d.N * sizeof(T);
    const int lst = std::min(128, ((d.N + 31) / 32) * 32);
    const size_t osmem = std::max((size_t)3 * d.N * sizeof(T), msmem);
    for (long long iter = 0; n > 0 && iter < cap; ++iter) {
      launch_dynamics(d.act, d.n_act, n);
      launch_kkt(d.act, d.n_act, n, method, hook != nullptr);
      launch_pcg(d.act, d.n_act, n, method, op.tol_lin, op.max_iter_lin);
      // structured path: the step recovery runs inside k_linesearch (unless a hook wants to see dz before the search)
      const int fuse_recover = (d.diag_mode && !legacy_ls && !hook) ? 1 : 0;
      if (!fuse_recover) launch_recover(d.act, d.n_act, n);
      if (hook) {
        B2T_CUDA(cudaStreamSynchronize(stream));
        if (hook(hook_user, B2T_HOOK_LINSYS, (int)iter)) return fail(B2T_ERR_INVALID, "iteration hook asked to stop");
      }
      if (!legacy_ls) {
        // k_linesearch also runs the outer (soft-constraint) update of the instances whose SQP loop exits in this pass
        // few active instances: all trials of a search at once (one block per instance and trial; bit-identical decisions)
        if (n <= ls_par_max && trials_exact) {
          Scope sc(this, B2T_K_TRIAL);
          if (d.integrator >= 2) k_linesearch_par<T, true><<<n * max_trials, lst, lsmem_par, stream>>>(d, op, 1, fuse_recover, max_trials, d_trial_out, d_trial_done);
          else k_linesearch_par<T><<<n * max_trials, lst, lsmem_par, stream>>>(d, op, 1, fuse_recover, max_trials, d_trial_out, d_trial_done);
          tick(B2T_K_TRIAL);
        } else if (d.integrator >= 2) { Scope sc(this, B2T_K_TRIAL); k_linesearch<T, true><<<n, lst, lsmem, stream>>>(d, op, 1, fuse_recover); tick(B2T_K_TRIAL); }
        else { Scope sc(this, B2T_K_TRIAL); k_linesearch<T><<<n, lst, lsmem, stream>>>(d, op, 1, fuse_recover); tick(B2T_K_TRIAL); }
      } else {
        { Scope sc(this, B2T_K_CTRL); k_iter_begin<T><<<cdiv(n, 128), 128, 0, stream>>>(d); tick(B2T_K_CTRL); }
        for (int t = 0; t < max_trials; ++t) {
          int* cur = (t % 2) ? d.ls_list1 : d.ls_list0;
          int* nxt = (t % 2) ? d.ls_list0 : d.ls_list1;
          if (d.integrator >= 2) { Scope sc(this, B2T_K_TRIAL); k_fd<T, true, true><<<cdiv((size_t)n * d.N, 128), 128, 0, stream>>>(d, cur, d.n_ls + t); tick(B2T_K_TRIAL); }
          else { Scope sc(this, B2T_K_TRIAL); k_fd<T, true><<<cdiv((size_t)n * d.N, 128), 128, 0, stream>>>(d, cur, d.n_ls + t); tick(B2T_K_TRIAL); }
          { Scope sc(this, B2T_K_MERIT); k_merit<T><<<n, mt, msmem, stream>>>(d, cur, d.n_ls + t, nxt, d.n_ls + t + 1, op); tick(B2T_K_MERIT); }
        }
        { Scope sc(this, B2T_K_CTRL); k_sqp_ctrl<T><<<cdiv(n, 128), 128, 0, stream>>>(d, op); tick(B2T_K_CTRL); }
        { Scope sc(this, B2T_K_CTRL); k_outer<T><<<n, mt, osmem, stream>>>(d, op, 0); tick(B2T_K_CTRL); }
        { Scope sc(this, B2T_K_MERIT); k_outer_begin<T><<<n, mt, msmem, stream>>>(d, d.restart_list, d.n_restart, op, 0); tick(B2T_K_MERIT); }
      }
      { Scope sc(this, B2T_K_CTRL); k_compact<T><<<1, 1024, 0, stream>>>(d, d_scratch, iter < PASS_TRACE_CAP ? d_pass_trace : nullptr, (int)iter); tick(B2T_K_CTRL); }
      n_passes = (int)iter + 1;
      if (lagged) {
        // no host round trip between passes: the count of pass i is copied asynchronously and read one pass later.  The active count
        // never grows, so the count of pass i-1 is a valid grid bound for pass i+1 (every kernel checks the device-side count), and
        // the pass launched after the last instance finished finds an empty list.
        const int slot = (int)(iter % COUNT_RING);
        B2T_CUDA(cudaMemcpyAsync(h_count + slot, d.n_act, sizeof(int), cudaMemcpyDeviceToHost, stream));
        B2T_CUDA(cudaEventRecord(ev_count[slot], stream));
        if (iter >= 1) {
          const int prev = (int)((iter - 1) % COUNT_RING);
          B2T_CUDA(cudaEventSynchronize(ev_count[prev]));
          n = h_count[prev];
        }
        continue;
      }
      B2T_CUDA(cudaMemcpyAsync(h_count, d.n_act, sizeof(int), cudaMemcpyDeviceToHost, stream));
      B2T_CUDA(cudaStreamSynchronize(stream));
      n = h_count[0];
      if (trace_active) fprintf(stderr, "%d ", n);
      if (hook && hook(hook_user, B2T_HOOK_STEP, (int)iter)) return fail(B2T_ERR_INVALID, "iteration hook asked to stop");
    }
    B2T_CUDA(cudaMemcpyAsync(h_count, d.n_act, sizeof(int), cudaMemcpyDeviceToHost, stream));
    B2T_CUDA(cudaEventRecord(ev1, stream));
    B2T_CUDA(cudaEventSynchronize(ev1));
    float ms = 0;
    B2T_CUDA(cudaEventElapsedTime(&ms, ev0, ev1));
    device_seconds = ms * 1e-3;
    collect_profile();
    B2T_CUDA(cudaGetLastError());
    // the pass budget (max_iter_soft * max_iter_SQP + 8) bounds every legal run; instances still active here mean a broken invariant
    if (h_count[0] > 0) return fail(B2T_ERR_UNFINISHED, std::to_string(h_count[0]) + " instances still active after the pass budget");
    return 0;
  }
  int get_pass_trace(int* counts, int cap, int* passes) override {
    if (passes) *passes = n_passes;
    if (counts && cap > 0) {
      B2T_CUDA(cudaSetDevice(device));
      const int n = std::min(std::min(cap, n_passes), (int)PASS_TRACE_CAP);
      B2T_CUDA(cudaMemcpy(counts, d_pass_trace, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost));
    }
    return 0;
  }
  const char* pcg_kernel_name() override {
    decide_pcg_variant();
    return pcg_variant == 8 ? "k_pcg_tm" : pcg_variant == 6 ? "k_pcg6" : pcg_variant == 5 ? "k_pcg4" : ((pcg_variant == 3 || pcg_variant == 4) ? "k_pcg3" : ((pcg_variant == 1 || pcg_variant == 2) ? "k_pcg2" : "k_pcg"));
  }

  int mpc_shift(const double* x_next, double* x0_out, double* u0_out, double* xnext_out) override {
    using namespace b2t;
    B2T_CUDA(cudaSetDevice(device));
    const size_t nb = (size_t)d.B * NX * sizeof(double);
    double* dxn = nullptr;
    if (x_next) {
      B2T_CUDA(cudaMemcpyAsync(stage_g, x_next, nb, cudaMemcpyHostToDevice, stream));
      dxn = stage_g;
    }
    double* o0 = stage_out; double* o1 = stage_out + (size_t)d.B * NX; double* o2 = o1 + (size_t)d.B * NU;
    if (d.integrator >= 2) k_mpc_shift<T, true><<<d.B, std::min(256, ((d.N + 31) / 32) * 32), 0, stream>>>(d, dxn, o0, o1, o2);
    else k_mpc_shift<T><<<d.B, std::min(256, ((d.N + 31) / 32) * 32), 0, stream>>>(d, dxn, o0, o1, o2);
    B2T_CUDA(cudaGetLastError());
    if (x0_out) B2T_CUDA(cudaMemcpyAsync(x0_out, o0, nb, cudaMemcpyDeviceToHost, stream));
    if (u0_out) B2T_CUDA(cudaMemcpyAsync(u0_out, o1, (size_t)d.B * NU * sizeof(double), cudaMemcpyDeviceToHost, stream));
    if (xnext_out) B2T_CUDA(cudaMemcpyAsync(xnext_out, o2, nb, cudaMemcpyDeviceToHost, stream));
    B2T_CUDA(cudaStreamSynchronize(stream));
    return 0;
  }

  // iLQR (oracle/ilqr.py is the specification): same work-list / outer-loop machinery as solve()
  int solve_ilqr(const b2t_options* o) override {
    using namespace b2t;
    if (!o) return fail(B2T_ERR_INVALID, "options required");
    if (d.N > ILQR_MAX_KNOTS) return fail(B2T_ERR_UNSUPPORTED, "iLQR supports at most 512 knot points");
    B2T_CUDA(cudaSetDevice(device));
    Opts<T> op = convert(o);
    launches = 0; device_seconds = 0;
    for (int i = 0; i < B2T_KERNEL_FAMILIES; ++i) { fam_seconds[i] = 0; fam_launches[i] = 0; }
    B2T_CUDA(cudaEventRecord(ev0, stream));
    const int B = d.B;
    if (d.hard.any) return fail(B2T_ERR_UNSUPPORTED, "iLQR supports soft limits only (README.md:17 of the reference says the same)");
    { int r = ensure_dense_gh(); if (r) return r; }
    const size_t msmem = (size_t)6 * d.N * sizeof(T);
    const int mt = merit_threads();
    k_init_state<T><<<cdiv(B, 128), 128, 0, stream>>>(d); tick(B2T_K_CTRL);
    { Scope sc(this, B2T_K_TRIAL); k_ilqr_rollout0<T><<<cdiv(B, 32), 32, 0, stream>>>(d); tick(B2T_K_TRIAL); }
    { Scope sc(this, B2T_K_MERIT); k_outer_begin<T><<<B, mt, msmem, stream>>>(d, d.act, d.n_act, op, 0); tick(B2T_K_MERIT); }
    B2T_CUDA(cudaGetLastError());
    int n = B;
    const long long cap = (long long)o->max_iter_soft * o->max_iter_SQP + 8;
    const size_t osmem = std::max((size_t)3 * d.N * sizeof(T), msmem);
    // block-parallel Riccati pass for nx >= 8 (arm6: 114.8 -> 10.1 ms / 2048 instances); one thread per instance for tiny systems
    // (cart-pole, nx = 4: 5.4 vs 9.8 ms).  B2T_ILQR_BW=1 / 2 forces a variant.
    // line search: all step lengths at once, ILQR_LPI lanes per instance (B2T_ILQR_SEARCH=1 selects the sequential kernel);
    // its trial buffer is allocated on first use
    int max_trials = 1;
    { double a = 1.0; while (a > o->alpha_min && max_trials < MAX_LS_TRIALS) { a *= o->alpha_factor; ++max_trials; } }
    {
      const char* se = getenv("B2T_ILQR_SEARCH");
      const bool sequential = (se && atoi(se) == 1) || max_trials > ILQR_LPI;
      if (sequential) ilqr_scratch = nullptr;
      else if (!ilqr_scratch_owned) {
        B2T_ALLOC(ilqr_scratch_owned, (size_t)(NM + 1) * d.K * ILQR_LPI);
        ilqr_scratch = ilqr_scratch_owned;
      } else ilqr_scratch = ilqr_scratch_owned;
    }
    const char* bwenv = getenv("B2T_ILQR_BW");
    const bool ilqr_bw_single = bwenv ? atoi(bwenv) == 1 : (NX < 8);
    for (long long iter = 0; n > 0 && iter < cap; ++iter) {
      launch_dynamics(d.act, d.n_act, n);
      { Scope sc(this, B2T_K_KKT); k_ilqr_cost<T><<<cdiv((size_t)n * d.N, 64), 64, 0, stream>>>(d, d.act, d.n_act); tick(B2T_K_KKT); }
      if (ilqr_bw_single) { Scope sc(this, B2T_K_SCHUR); k_ilqr_backward<T><<<cdiv(n, 32), 32, 0, stream>>>(d, d.act, d.n_act); tick(B2T_K_SCHUR); }
      else { Scope sc(this, B2T_K_SCHUR); k_ilqr_backward2<T><<<n, ILQR_BW_THREADS, 0, stream>>>(d, d.act, d.n_act); tick(B2T_K_SCHUR); }
      if (ilqr_scratch) { Scope sc(this, B2T_K_TRIAL); k_ilqr_search2<T><<<cdiv((size_t)n * ILQR_LPI, 64), 64, 0, stream>>>(d, op, ilqr_scratch, max_trials); tick(B2T_K_TRIAL); }
      else { Scope sc(this, B2T_K_TRIAL); k_ilqr_search<T><<<cdiv(n, 32), 32, 0, stream>>>(d, op); tick(B2T_K_TRIAL); }
      { Scope sc(this, B2T_K_CTRL); k_outer<T><<<n, mt, osmem, stream>>>(d, op, 1); tick(B2T_K_CTRL); }
      { Scope sc(this, B2T_K_CTRL); k_compact<T><<<1, 1024, 0, stream>>>(d, d_scratch); tick(B2T_K_CTRL); }
      B2T_CUDA(cudaMemcpyAsync(h_count, d.n_act, sizeof(int), cudaMemcpyDeviceToHost, stream));
      B2T_CUDA(cudaStreamSynchronize(stream));
      n = h_count[0];
    }
    B2T_CUDA(cudaEventRecord(ev1, stream));
    B2T_CUDA(cudaEventSynchronize(ev1));
    float ms = 0;
    B2T_CUDA(cudaEventElapsedTime(&ms, ev0, ev1));
    device_seconds = ms * 1e-3;
    collect_profile();
    B2T_CUDA(cudaGetLastError());
    return 0;
  }

  int solve_host(const double* x0, const double* u0, const double* xg, int method, const b2t_options* o, double* xo, double* uo,
                 int* st) override {
    int r;
    if ((r = set_trajectory(x0, u0, 0))) return r;
    if (xg && (r = set_goals(xg, 0))) return r;
    if ((r = solve(method, o))) return r;
    if ((r = get_trajectory(xo, uo, 0))) return r;
    if (st && (r = get_status(st))) return r;
    return 0;
  }

  // ------------------------------------------------------------------ stages (parity tests)
  int all_list() {   // act = 0..B-1
    using namespace b2t;
    k_init_state<T><<<cdiv(d.B, 128), 128, 0, stream>>>(d);
    B2T_CUDA(cudaGetLastError());
    return 0;
  }
  int stage_dynamics() override {
    B2T_CUDA(cudaSetDevice(device));
    int r = all_list(); if (r) return r;
    B2T_CUDA(cudaMemsetAsync(d.dyn_ok, 0, (size_t)d.B * sizeof(int), stream));
    for (int i = 0; i < B2T_KERNEL_FAMILIES; ++i) { fam_seconds[i] = 0; fam_launches[i] = 0; }
    launch_dynamics(d.act, d.n_act, d.B);
    B2T_CUDA(cudaGetLastError());
    B2T_CUDA(cudaStreamSynchronize(stream));
    collect_profile();           // with profiling on: CUDA-event times of k_fd / k_fd_grad (b2t_get_kernel_times)
    return 0;
  }
  int stage_kkt(double rho, int method) override {
    using namespace b2t;
    B2T_CUDA(cudaSetDevice(device));
    explicit_system = false;
    int r = all_list(); if (r) return r;
    k_fill<T><<<cdiv(d.B, 128), 128, 0, stream>>>(d.rho, (size_t)d.B, (T)rho);
    launch_dynamics(d.act, d.n_act, d.B);
    launch_kkt(d.act, d.n_act, d.B, method, true);
    B2T_CUDA(cudaGetLastError());
    B2T_CUDA(cudaStreamSynchronize(stream));
    return 0;
  }
  int stage_pcg(int method, double tol, int max_iter, int* iters) override {
    B2T_CUDA(cudaSetDevice(device));
    launch_pcg(d.act, d.n_act, d.B, method, (T)tol, max_iter);
    B2T_CUDA(cudaGetLastError());
    if (iters) B2T_CUDA(cudaMemcpyAsync(iters, d.pcg_iters, (size_t)d.B * sizeof(int), cudaMemcpyDeviceToHost, stream));
    B2T_CUDA(cudaStreamSynchronize(stream));
    return 0;
  }
  int set_block_system(const double* Sd, const double* So, const double* gam) override {
    using namespace b2t;
    if (!Sd || !So || !gam) return fail(B2T_ERR_INVALID, "Sd, So, gamma required");
    B2T_CUDA(cudaSetDevice(device));
    const size_t nb = d.K * NX * NX, nv = d.K * NX;
    if ((2 * nb + nv) * sizeof(double) > stage_out_bytes) return fail(B2T_ERR_INVALID, "system too large for staging");
    double* p0 = stage_out; double* p1 = p0 + nb; double* p2 = p1 + nb;
    B2T_CUDA(cudaMemcpyAsync(p0, Sd, nb * sizeof(double), cudaMemcpyHostToDevice, stream));
    B2T_CUDA(cudaMemcpyAsync(p1, So, nb * sizeof(double), cudaMemcpyHostToDevice, stream));
    B2T_CUDA(cudaMemcpyAsync(p2, gam, nv * sizeof(double), cudaMemcpyHostToDevice, stream));
    int r = all_list(); if (r) return r;
    explicit_system = true;
    k_set_block_system<T><<<cdiv(d.K, 128), 128, 0, stream>>>(d, p0, p1, p2);
    B2T_CUDA(cudaGetLastError());
    B2T_CUDA(cudaStreamSynchronize(stream));
    return 0;
  }
  int stage_precond(int method) override {
    using namespace b2t;
    B2T_CUDA(cudaSetDevice(device));
    k_pinv<T><<<cdiv(d.K, 128), 128, 0, stream>>>(d, d.act, d.n_act, method == B2T_METHOD_PCG_J ? 1 : 0);
    B2T_CUDA(cudaGetLastError());
    B2T_CUDA(cudaStreamSynchronize(stream));
    return 0;
  }
  int stage_recover() override {
    B2T_CUDA(cudaSetDevice(device));
    launch_recover(d.act, d.n_act, d.B);
    B2T_CUDA(cudaGetLastError());
    B2T_CUDA(cudaStreamSynchronize(stream));
    return 0;
  }
  int stage_merit(double alpha, double* J, double* c, double* D) override {
    using namespace b2t;
    B2T_CUDA(cudaSetDevice(device));
    k_fill<T><<<cdiv(d.B, 128), 128, 0, stream>>>(d.alpha, (size_t)d.B, (T)alpha);
    if (d.integrator >= 2) k_fd<T, true, true><<<cdiv(d.K, 128), 128, 0, stream>>>(d, d.act, d.n_act);
    else k_fd<T, true><<<cdiv(d.K, 128), 128, 0, stream>>>(d, d.act, d.n_act);
    k_merit_only<T><<<d.B, merit_threads(), (size_t)6 * d.N * sizeof(T), stream>>>(d, d.J, d.c, d.D);
    B2T_CUDA(cudaGetLastError());
    std::vector<T> h(d.B);
    T* src[3] = {d.J, d.c, d.D};
    double* dst[3] = {J, c, D};
    for (int i = 0; i < 3; ++i) {
      if (!dst[i]) continue;
      B2T_CUDA(cudaMemcpy(h.data(), src[i], (size_t)d.B * sizeof(T), cudaMemcpyDeviceToHost));
      for (int b = 0; b < d.B; ++b) dst[i][b] = (double)h[b];
    }
    return 0;
  }
  int fetch(int which, double* out) override {
    using namespace b2t;
    if (!out) return fail(B2T_ERR_INVALID, "out required");
    B2T_CUDA(cudaSetDevice(device));
    const size_t K = d.K;
    const T* src = nullptr; int E = 0;
    switch (which) {
      case B2T_ARR_X: src = d.x; E = NX; break;
      case B2T_ARR_U: src = d.u; E = NU; break;
      case B2T_ARR_XKP1: src = d.xkp1; E = NX; break;
      case B2T_ARR_DQDD:
        if (d.integrator >= 2) return fail(B2T_ERR_UNSUPPORTED, "integrator types 2 / 3 evaluate the dynamics gradient per stage: fetch B2T_ARR_AB instead");
        src = d.dyn; E = NDYN; break;
      case B2T_ARR_GHAT: src = d.Gh; E = NM * NM; break;
      case B2T_ARR_G: src = d.g; E = NM; break;
      case B2T_ARR_DZ: src = d.dz; E = NM; break;
      case B2T_ARR_XN: src = d.xn; E = NX; break;
      case B2T_ARR_UN: src = d.un; E = NU; break;
      case B2T_ARR_SD: src = d.Sd; E = NX * NX; break;
      case B2T_ARR_SO: src = d.So; E = NX * NX; break;
      case B2T_ARR_PD: src = d.Pd; E = NX * NX; break;
      case B2T_ARR_GAMMA: src = d.gam; E = NX; break;
      case B2T_ARR_L: src = d.l; E = NX; break;
      case B2T_ARR_COST_VALUE: E = 1; break;
      case B2T_ARR_COST_GRAD: E = NM; break;
      case B2T_ARR_COST_HESS: E = NM * NM; break;
      case B2T_ARR_COST_ERR: E = NX; break;
      case B2T_ARR_KKT_HESS: E = NM * NM; break;
      case B2T_ARR_AB: E = NX * NM; break;
      case B2T_ARR_SOFT_VALUE: E = 1; break;
      case B2T_ARR_SOFT_GRAD: E = NM; break;
      case B2T_ARR_COST_JTOT: E = NX * NX; break;
      case B2T_ARR_PLANT_TERMS: E = PLANT_TERMS; break;
      case B2T_ARR_NU_TRACE: {
        std::vector<T> h((size_t)d.B * NU_TRACE_LEN);
        B2T_CUDA(cudaStreamSynchronize(stream));
        B2T_CUDA(cudaMemcpy(h.data(), d.nu_trace, h.size() * sizeof(T), cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < h.size(); ++i) out[i] = (double)h[i];
        return 0;
      }
      default: return fail(B2T_ERR_INVALID, "unknown array id");
    }
    const size_t n = K * E;
    if (n * sizeof(double) > stage_out_bytes) return fail(B2T_ERR_INVALID, "array too large for staging");
    if (which >= B2T_ARR_COST_VALUE && which <= B2T_ARR_SOFT_GRAD) k_cost_eval<T><<<cdiv(K, 64), 64, 0, stream>>>(d, which - B2T_ARR_COST_VALUE, stage_out);
    else if (which == B2T_ARR_COST_JTOT) k_cost_eval<T><<<cdiv(K, 64), 64, 0, stream>>>(d, 8, stage_out);
    else if (which == B2T_ARR_PLANT_TERMS) k_plant_eval<T><<<cdiv(K, 64), 64, 0, stream>>>(d, stage_out);
    else if (which == B2T_ARR_GHAT && d.diag_mode) k_fetch_ghat_diag<T><<<cdiv(K, 128), 128, 0, stream>>>(d, stage_out);
    else k_fetch_soa<T><<<cdiv(K, 128), 128, 0, stream>>>(src, K, E, stage_out);
    B2T_CUDA(cudaGetLastError());
    B2T_CUDA(cudaMemcpyAsync(out, stage_out, n * sizeof(double), cudaMemcpyDeviceToHost, stream));
    B2T_CUDA(cudaStreamSynchronize(stream));
    return 0;
  }
};
}  // namespace

template <typename T>
__global__ void __launch_bounds__(256) k_fma_peak(T* out, int iters, T a, T b) {
  T acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = (T)(threadIdx.x + i);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = acc[i] * a + b;
  }
  T s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += acc[i];
  if (s == (T)123456789) out[0] = s;     // never true: keeps the chain alive
}
template <typename T>
int fma_peak(int device, double* tflops) {
  B2T_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  B2T_CUDA(cudaGetDeviceProperties(&prop, device));
  T* out = nullptr;
  B2T_CUDA(cudaMalloc(&out, sizeof(T)));
  cudaEvent_t e0, e1;
  B2T_CUDA(cudaEventCreate(&e0)); B2T_CUDA(cudaEventCreate(&e1));
  const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 4096;
  double best = 0;
  for (int rep = 0; rep < 6; ++rep) {
    B2T_CUDA(cudaEventRecord(e0));
    k_fma_peak<T><<<blocks, threads>>>(out, iters, (T)0.999999, (T)1e-9);
    B2T_CUDA(cudaEventRecord(e1));
    B2T_CUDA(cudaEventSynchronize(e1));
    float ms = 0;
    B2T_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    const double flops = 2.0 * 16 * (double)iters * threads * blocks;
    if (rep > 0) best = std::max(best, flops / (ms * 1e-3) / 1e12);
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(out);
  *tflops = best;
  return 0;
}

#if B2T_PART != 2
SolverBase* b2t_make_solver_f64() { return new SolverT<double>(); }
int b2t_fma_peak_f64(int device, double* tflops) { return fma_peak<double>(device, tflops); }
#endif
#if B2T_PART != 1
SolverBase* b2t_make_solver_f32() { return new SolverT<float>(); }
int b2t_fma_peak_f32(int device, double* tflops) { return fma_peak<float>(device, tflops); }
#endif

#if B2T_PART != 2
struct b2t_solver { SolverBase* impl; };

extern "C" {
int b2t_abi_version(void) { return B2T_ABI_VERSION; }
const char* b2t_model_name(void) { return B2T_MODEL_NAME; }
const char* b2t_model_digest(void) { return B2T_MODEL_DIGEST; }
int b2t_model_dims(int* nq, int* nx, int* nu) {
  if (nq) *nq = b2t::NQ;
  if (nx) *nx = b2t::NX;
  if (nu) *nu = b2t::NU;
  return 0;
}
const char* b2t_last_error(void) { return b2t_err_slot().c_str(); }
void b2t_default_options(b2t_options* o) {
  if (!o) return;
  o->exit_tolerance_linSys = 1e-6; o->max_iter_linSys = 100;
  o->exit_tolerance_SQP = 1e-6; o->max_iter_SQP = 100;
  o->alpha_factor = 0.5; o->alpha_min = 0.005;
  o->rho_factor = 4; o->rho_min = 1e-3; o->rho_max = 1e3; o->rho_init = 1e-3;
  o->expected_reduction_min = 0.05; o->expected_reduction_max = 3;
  o->exit_tolerance_soft = 1e-6; o->max_iter_soft = 10;
  o->merit_mu = 10;
}
int b2t_solver_create(const b2t_problem_desc* desc, int device, b2t_solver** out) {
  if (!desc || !out) return fail(B2T_ERR_INVALID, "desc and out required");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(B2T_ERR_CUDA, "no CUDA device: this library has no CPU fallback");
  if (device < 0 || device >= ndev) return fail(B2T_ERR_INVALID, "bad device index");
  SolverBase* impl = nullptr;
  if (desc->dtype == B2T_F64) impl = b2t_make_solver_f64();
  else if (desc->dtype == B2T_F32) impl = b2t_make_solver_f32();
  else return fail(B2T_ERR_INVALID, "dtype");
  int r = impl->init(desc, device);
  if (r) { delete impl; return r; }
  *out = new b2t_solver{impl};
  return 0;
}
int b2t_solver_destroy(b2t_solver* s) {
  if (!s) return 0;
  delete s->impl;
  delete s;
  return 0;
}
size_t b2t_workspace_bytes(const b2t_solver* s) { return s ? s->impl->ws_bytes : 0; }
int b2t_set_stream(b2t_solver* s, void* st) { if (!s) return fail(B2T_ERR_INVALID, "null"); s->impl->stream = (cudaStream_t)st; return 0; }
#define B2T_FWD(call) do { if (!s) return fail(B2T_ERR_INVALID, "null solver"); return s->impl->call; } while (0)
int b2t_set_trajectory(b2t_solver* s, const double* x, const double* u, int od) { B2T_FWD(set_trajectory(x, u, od)); }
int b2t_set_goals(b2t_solver* s, const double* xg, int od) { B2T_FWD(set_goals(xg, od)); }
int b2t_set_initial_state(b2t_solver* s, const double* xs) { B2T_FWD(set_initial_state(xs)); }
int b2t_set_multipliers(b2t_solver* s, const double* mu, const double* lam, const double* phi) { B2T_FWD(set_multipliers(mu, lam, phi)); }
int b2t_reset_multipliers(b2t_solver* s) { B2T_FWD(reset_multipliers()); }
int b2t_sqp_solve(b2t_solver* s, int method, const b2t_options* o) { B2T_FWD(solve(method, o)); }
int b2t_set_iteration_hook(b2t_solver* s, b2t_iteration_hook hook, void* user) {
  if (!s) return fail(B2T_ERR_INVALID, "null solver");
  s->impl->hook = hook; s->impl->hook_user = user;
  return 0;
}
int b2t_ilqr_solve(b2t_solver* s, const b2t_options* o) { B2T_FWD(solve_ilqr(o)); }
int b2t_mpc_shift(b2t_solver* s, const double* xn, double* x0, double* u0, double* xno) { B2T_FWD(mpc_shift(xn, x0, u0, xno)); }
int b2t_get_trajectory(b2t_solver* s, double* x, double* u, int od) { B2T_FWD(get_trajectory(x, u, od)); }
int b2t_get_status(b2t_solver* s, int* st) { B2T_FWD(get_status(st)); }
int b2t_get_scalars(b2t_solver* s, double* sc) { B2T_FWD(get_scalars(sc)); }
int b2t_get_trace(b2t_solver* s, double* tr, int cap) { B2T_FWD(get_trace(tr, cap)); }
int b2t_get_multipliers(b2t_solver* s, double* mu, double* lam, double* phi) { B2T_FWD(get_multipliers(mu, lam, phi)); }
int b2t_get_launch_stats(b2t_solver* s, long long* l, double* sec) {
  if (!s) return fail(B2T_ERR_INVALID, "null solver");
  if (l) *l = s->impl->launches;
  if (sec) *sec = s->impl->device_seconds;
  return 0;
}
int b2t_set_profiling(b2t_solver* s, int mode) {
  if (!s) return fail(B2T_ERR_INVALID, "null solver");
  if (mode < 0 || mode >= 2 + B2T_KERNEL_FAMILIES) return fail(B2T_ERR_INVALID, "profiling mode");
  s->impl->profiling = mode != 0;
  s->impl->profile_mask = mode == 0 ? 0u : (mode == 1 ? (1u << B2T_KERNEL_FAMILIES) - 1u : 1u << (mode - 2));
  return 0;
}
int b2t_get_kernel_times(b2t_solver* s, double* sec, long long* l) {
  if (!s) return fail(B2T_ERR_INVALID, "null solver");
  for (int i = 0; i < B2T_KERNEL_FAMILIES; ++i) { if (sec) sec[i] = s->impl->fam_seconds[i]; if (l) l[i] = s->impl->fam_launches[i]; }
  return 0;
}
int b2t_sqp_solve_host(b2t_solver* s, const double* x0, const double* u0, const double* xg, int method, const b2t_options* o,
                       double* xo, double* uo, int* st) { B2T_FWD(solve_host(x0, u0, xg, method, o, xo, uo, st)); }
int b2t_stage_dynamics(b2t_solver* s) { B2T_FWD(stage_dynamics()); }
int b2t_stage_kkt(b2t_solver* s, double rho, int method) { B2T_FWD(stage_kkt(rho, method)); }
int b2t_stage_pcg(b2t_solver* s, int method, double tol, int mi, int* it) { B2T_FWD(stage_pcg(method, tol, mi, it)); }
int b2t_stage_recover(b2t_solver* s) { B2T_FWD(stage_recover()); }
int b2t_set_block_system(b2t_solver* s, const double* a, const double* b, const double* c) { B2T_FWD(set_block_system(a, b, c)); }
int b2t_stage_precond(b2t_solver* s, int method) { B2T_FWD(stage_precond(method)); }
int b2t_stage_merit(b2t_solver* s, double a, double* J, double* c, double* D) { B2T_FWD(stage_merit(a, J, c, D)); }
int b2t_fetch(b2t_solver* s, int which, double* out) { B2T_FWD(fetch(which, out)); }
int b2t_get_pass_trace(b2t_solver* s, int* counts, int cap, int* passes) { B2T_FWD(get_pass_trace(counts, cap, passes)); }
const char* b2t_pcg_kernel_name(b2t_solver* s) { return s ? s->impl->pcg_kernel_name() : ""; }
int b2t_measure_fma_peak(int device, int dtype, double* tflops) {
  if (!tflops) return fail(B2T_ERR_INVALID, "tflops required");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return fail(B2T_ERR_CUDA, "no such CUDA device");
  return dtype == B2T_F32 ? b2t_fma_peak_f32(device, tflops) : b2t_fma_peak_f64(device, tflops);
}
}
#endif  // B2T_PART != 2
