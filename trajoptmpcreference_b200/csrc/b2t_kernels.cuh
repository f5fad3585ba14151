// b2t_kernels.cuh -- CUDA kernels of the batched SQP / Schur / GBD-PCG path (sm_100a).
//
// Data layout in HBM (K = batch * N knot points, t = b * N + k):
//   per-knot arrays are element-major ("SoA over knots"):  a[e][t]  -> thread t of a per-knot kernel reads a
//   coalesced 8-byte stream for every element e;
//   the block-tridiagonal Schur complement and preconditioner blocks use the same layout, one nx*nx block per knot:
//   Sd[i*nx+c][t] = S_kk[i][c], So[i*nx+c][t] = S_{k,k-1}[i][c], Pd likewise; gamma[i][t], l[i][t].  The per-knot
//   assembly kernels write them fully coalesced; the per-instance PCG block loads them once into registers / shared memory.
// Control flow (outer AL loop, SQP loop, line search) lives on the device in per-instance state words; the host
// only launches the fixed kernel sequence and reads one counter per SQP iteration.
#pragma once
#include <cuda_runtime.h>
#include "b2t_core.cuh"

namespace b2t {

enum Phase { PH_SQP = 0, PH_OUTER = 1, PH_DONE = 2 };
enum { TRACE_FIELDS = 12 };
enum { MAX_LS_TRIALS = 32 };
enum { NU_TRACE_LEN = 128 };

template <typename T>
struct Opts {
  T tol_lin; int max_iter_lin;
  T tol_sqp; int max_iter_sqp;
  T alpha_factor, alpha_min;
  T rho_factor, rho_min, rho_max, rho_init;
  T er_min, er_max;
  T tol_soft; int max_iter_soft;
  T merit_mu;
};

template <typename T>
struct Dev {
  int B, N, integrator;
  int diag_mode;      // 1: Q, QF, R diagonal -> Ghat_k kept as diag(dinv) - s h h^T (Sherman-Morrison), structured Schur kernels
  T dt, gravity;
  size_t K;
  // per-knot SoA
  T *x, *u, *xn, *un, *xkp1, *xkp1n, *dyn, *vaf, *Gh, *g, *Gg, *dz, *mu, *lam, *phi;
  T *ABf;             // [nx*nm][K] full [A_k B_k] of integrator types 2 / 3 (k_ab_multi); null for types 0 / 1, whose [A B] is rebuilt from dyn
  // block-tridiagonal system, per-knot SoA: [nx*nx][K] and [nx][K]
  T *Sd, *So, *Pd, *gam, *l;
  // per-instance
  T *xs, *xg;
  CostParams<T> cost;
  LimitParams<T> lim;     // soft limits (penalty / augmented Lagrangian)
  LimitParams<T> hard;    // hard ACTIVE_SET limits (mode[i] != 0), exact methods only
  T mu_factor[3], mu_max[3], phi_factor[3], mu_init[3], phi_init[3];
  T *rho, *drho, *J, *c, *merit, *alpha, *deltaJ, *D, *ratio;
  int *ls_iter, *sqp_iter, *outer_iter, *exit_sqp, *exit_soft, *phase, *err, *pcg_iters, *tot_qp, *tot_pcg, *tot_trials;
  int *act, *n_act;
  int *ls_list0, *ls_list1;
  int *restart_list, *n_restart;
  int *dyn_ok;        // [B] 1: dynamics, gradient and x+ of the current (x, u) are still valid (the last line search failed: x, u unchanged)
  int *n_ls;          // [MAX_LS_TRIALS + 1]
  T* nu_trace;        // optional [B][NU_TRACE_LEN]: |r^T Pinv r| of every PCG iteration (PCG.pcg's `trace`, PCG.py:82,95); null = off
  T* trace;           // [B][trace_cap][TRACE_FIELDS]
  int* trace_rows;
  int trace_cap;
};

// [A_k B_k] of knot t (NX x NM row-major): rebuilt from dqdd for the Euler integrators, loaded for the multi-stage ones
template <typename T>
__device__ __forceinline__ void load_AB(const Dev<T>& d, size_t t, T* AB) {
  if (d.integrator >= 2) {
    for (int i = 0; i < NX * NM; ++i) AB[i] = d.ABf[(size_t)i * d.K + t];
  } else {
    T dq[NDYN];
    for (int i = 0; i < NDYN; ++i) dq[i] = d.dyn[(size_t)i * d.K + t];
    build_AB(d.integrator, dq, d.dt, AB);
  }
}

template <typename T>
__device__ __forceinline__ void load_xu(const T* xs_, const T* us_, size_t K, size_t t, bool terminal, T* x, T* u) {
  for (int i = 0; i < NX; ++i) x[i] = xs_[(size_t)i * K + t];
  for (int i = 0; i < NU; ++i) u[i] = terminal ? T(0) : us_[(size_t)i * K + t];
}

// -----------------------------------------------------------------------------------------------------------------
// k_fd: forward dynamics + integrator step of one knot per thread.
//   TRIAL = false: reads (x,u); writes xkp1, Minv block of dyn, v/a/f of rnea(q,qd,qdd) (inputs of k_fd_grad)
//   TRIAL = true : first forms the trial point xn = x - alpha dz_x, un = u - alpha dz_u (SQP :617-622), then xkp1n
// -----------------------------------------------------------------------------------------------------------------
template <typename T, bool TRIAL, bool MS = false>      // MS: multi-stage integrator (types 2 / 3), its own instantiation so that the Euler code is untouched
__global__ void __launch_bounds__(128) k_fd(Dev<T> d, const int* list, const int* count) {
  const size_t gt = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int slot = (int)(gt / d.N);
  if (slot >= *count) return;
  const int k = (int)(gt % d.N);
  const int b = list[slot];
  if constexpr (!TRIAL) { if (d.dyn_ok[b]) return; }      // unchanged iterate (failed line search): nothing to recompute
  const size_t t = (size_t)b * d.N + k;
  const size_t K = d.K;
  const bool terminal = (k == d.N - 1);
  T x[NX], u[NU];
  load_xu(d.x, d.u, K, t, terminal, x, u);
  if constexpr (TRIAL) {
    const T al = d.alpha[b];
    for (int i = 0; i < NX; ++i) { x[i] = x[i] - al * d.dz[(size_t)i * K + t]; d.xn[(size_t)i * K + t] = x[i]; }
    if (!terminal)
      for (int i = 0; i < NU; ++i) { u[i] = u[i] - al * d.dz[(size_t)(NX + i) * K + t]; d.un[(size_t)i * K + t] = u[i]; }
  }
  if (terminal) return;
  T qdd[NJ], Minv[NJ * NJ], v[NJ][6], a[NJ][6], f[NJ][6], xnext[NX];
  T* outp = TRIAL ? d.xkp1n : d.xkp1;
  if constexpr (MS) {      // midpoint / rk3 (the reference's variants): x+ here, [A B] in k_ab_multi
    integrator_multi_value(d.integrator, x, u, d.gravity, d.dt, xnext);
    for (int i = 0; i < NX; ++i) outp[(size_t)i * K + t] = xnext[i];
    return;
  }
  forward_dynamics<T, !TRIAL>(x, x + NJ, u, d.gravity, qdd, Minv, v, a, f);
  integrate(d.integrator, x, qdd, d.dt, xnext);
  for (int i = 0; i < NX; ++i) outp[(size_t)i * K + t] = xnext[i];
  if constexpr (!TRIAL) {
    for (int i = 0; i < NJ; ++i)
      for (int j = 0; j < NJ; ++j) d.dyn[(size_t)(i * 3 * NJ + 2 * NJ + j) * K + t] = Minv[i * NJ + j];
    for (int j = 0; j < NJ; ++j)
      for (int i = 0; i < 6; ++i) {
        d.vaf[(size_t)(j * 6 + i) * K + t] = v[j][i];
        d.vaf[(size_t)(6 * NJ + j * 6 + i) * K + t] = a[j][i];
        d.vaf[(size_t)(12 * NJ + j * 6 + i) * K + t] = f[j][i];
      }
  }
}

// -----------------------------------------------------------------------------------------------------------------
// k_fd_grad: one column of dqdd/d(q,qd) per thread.  Thread index = col-major over (column, slot, knot) so that a warp
// shares the column (uniform branches) and reads consecutive knots (coalesced).
// -----------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(128) k_fd_grad(Dev<T> d, const int* list, const int* count) {
  // 1-D grid, the column index fastest: the 2 NJ blocks that differentiate the SAME 128 knots are scheduled together, so v, a, f, Minv
  // are read from DRAM once and served from L2 to the other columns (with the column as grid.y every column pass streamed the whole
  // input again: 996 MB instead of 158 MB of DRAM reads per launch at 116 k knots)
  const int colid = (int)(blockIdx.x % (2 * NJ));   // 0 .. 2*NJ-1
  const size_t gt = (size_t)(blockIdx.x / (2 * NJ)) * blockDim.x + threadIdx.x;
  const int slot = (int)(gt / d.N);
  if (slot >= *count) return;
  const int k = (int)(gt % d.N);
  if (k == d.N - 1) return;
  const bool is_qd = colid >= NJ;
  const int col = is_qd ? colid - NJ : colid;
  const int b = list[slot];
  if (d.dyn_ok[b]) return;
  const size_t t = (size_t)b * d.N + k;
  const size_t K = d.K;
  T q[NJ], qd[NJ], Minv[NJ * NJ], v[NJ][6], a[NJ][6], f[NJ][6];
  for (int i = 0; i < NJ; ++i) { q[i] = d.x[(size_t)i * K + t]; qd[i] = d.x[(size_t)(NJ + i) * K + t]; }
  for (int i = 0; i < NJ; ++i)
    for (int j = 0; j < NJ; ++j) Minv[i * NJ + j] = d.dyn[(size_t)(i * 3 * NJ + 2 * NJ + j) * K + t];
  for (int j = 0; j < NJ; ++j)
    for (int i = 0; i < 6; ++i) {
      v[j][i] = d.vaf[(size_t)(j * 6 + i) * K + t];
      a[j][i] = d.vaf[(size_t)(6 * NJ + j * 6 + i) * K + t];
      f[j][i] = d.vaf[(size_t)(12 * NJ + j * 6 + i) * K + t];
    }
  T out[NJ];
  fd_grad_column(q, qd, v, a, f, Minv, d.gravity, col, is_qd, out);
  for (int i = 0; i < NJ; ++i) d.dyn[(size_t)(i * 3 * NJ + colid) * K + t] = out[i];
}

// -----------------------------------------------------------------------------------------------------------------
// k_ab_multi: [A_k B_k] of integrator types 2 / 3 (TrajoptPlant.py:140-205 as written there), one knot per thread, into ABf.
// -----------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(64) k_ab_multi(Dev<T> d, const int* list, const int* count) {
  const size_t gt = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int slot = (int)(gt / d.N);
  if (slot >= *count) return;
  const int k = (int)(gt % d.N);
  if (k == d.N - 1) return;
  const int b = list[slot];
  if (d.dyn_ok[b]) return;
  const size_t t = (size_t)b * d.N + k;
  T x[NX], u[NU], AB[NX * NM];
  load_xu(d.x, d.u, d.K, t, false, x, u);
  integrator_multi_AB(d.integrator, x, u, d.gravity, d.dt, AB);
  for (int i = 0; i < NX * NM; ++i) d.ABf[(size_t)i * d.K + t] = AB[i];
}

// -----------------------------------------------------------------------------------------------------------------
// k_kkt: per knot  G_k = hess + gck gck^T + rho I,  Ghat_k = G_k^-1,  g_k = grad + gck,  Gg_k = Ghat_k g_k
// (formKKTSystemBlocks :216-225,252-260; solveKKTSystem_Schur :419-422)
// -----------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(64) k_kkt(Dev<T> d, const int* list, const int* count) {
  const size_t gt = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int slot = (int)(gt / d.N);
  if (slot >= *count) return;
  const int k = (int)(gt % d.N);
  const int b = list[slot];
  const size_t t = (size_t)b * d.N + k;
  const size_t K = d.K;
  const bool terminal = (k == d.N - 1);
  T z[NM], xg[NX];
  load_xu(d.x, d.u, K, t, terminal, z, z + NX);
  for (int i = 0; i < NX; ++i) xg[i] = d.xg[(size_t)i * d.B + b];
  T g[NM], G[NM * NM];
  cost_grad_hess<T, true>(d.cost, z, z + NX, xg, k, terminal, g, G);
  if (d.lim.any) {
    T gck[NM];
    soft_grad(d.lim, z, d.mu + t, d.lam + t, K, terminal, gck);
    for (int i = 0; i < NM; ++i) g[i] += gck[i];
    for (int i = 0; i < NM; ++i)
      for (int j = 0; j < NM; ++j) G[i * NM + j] += gck[i] * gck[j];
  }
  const T rho = d.rho[b];
  const int M = terminal ? NX : NM;
  for (int i = 0; i < NM; ++i) G[i * NM + i] += rho;
  // Hard ACTIVE_SET rows fix the step of their coordinate, dz_A = t.  Eliminating them from the KKT system leaves the same
  // block-tridiagonal Schur system with  Ghat' = E_F (G_FF)^-1 E_F^T,  g' = g - G[:, A] t  and  Ghat' g' + E_A t  in place of Ghat g:
  //   dz = Ghat' (g' - C^T l) + E_A t,   S = -C Ghat' C^T,   gamma = c - C (Ghat' g' + E_A t)
  // (the solution of the reference's enlarged KKT system, TrajoptMPCReference.py:238-248, :313-359)
  T tfix[NM];
  bool act[NM];
  bool any_act = false;
  for (int i = 0; i < NM; ++i) { tfix[i] = T(0); act[i] = false; }
  if (d.hard.any)
    for (int i = 0; i < M; ++i) { act[i] = hard_active(d.hard, z, i, &tfix[i]); any_act = any_act || act[i]; }
  if (any_act) {
    for (int a = 0; a < M; ++a) {
      if (!act[a]) continue;
      for (int j = 0; j < M; ++j)
        if (!act[j]) g[j] -= G[j * NM + a] * tfix[a];
    }
    for (int a = 0; a < M; ++a) {
      if (!act[a]) continue;
      for (int j = 0; j < M; ++j) { G[a * NM + j] = T(0); G[j * NM + a] = T(0); }
      G[a * NM + a] = T(1);
      g[a] = T(0);
    }
  }
  spd_inverse_inplace(G, M, NM);
  if (any_act)
    for (int a = 0; a < M; ++a)
      if (act[a]) G[a * NM + a] = T(0);
  if (terminal)
    for (int i = 0; i < NM; ++i)
      for (int j = 0; j < NM; ++j)
        if (i >= NX || j >= NX) G[i * NM + j] = T(0);
  for (int i = 0; i < NM; ++i) {
    T acc = T(0);
    for (int j = 0; j < M; ++j) acc += G[i * NM + j] * g[j];
    d.Gg[(size_t)i * K + t] = (i < M) ? acc + tfix[i] : T(0);
    d.g[(size_t)i * K + t] = (i < M) ? g[i] : T(0);
  }
  for (int i = 0; i < NM * NM; ++i) d.Gh[(size_t)i * K + t] = G[i];
}

// -----------------------------------------------------------------------------------------------------------------
// k_schur: per block row j  S_jj, S_j,j-1, gamma_j and the preconditioner diagonal block (block form of
// S = -C Ghat C^T, gamma = c - C Ghat g, :423-424; PCG.compute_preconditioner PCG.py:166-212).
//   precond: 0 = Jacobi (diag(S)^-1), 1 = block Jacobi / symmetric stair (S_jj^-1)
// -----------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(64) k_schur(Dev<T> d, const int* list, const int* count, int jacobi) {
  const size_t gt = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int slot = (int)(gt / d.N);
  if (slot >= *count) return;
  const int j = (int)(gt % d.N);
  const int b = list[slot];
  const size_t t = (size_t)b * d.N + j;
  const size_t K = d.K;
  (void)0;
  T Sd[NX * NX], gam[NX];
  T* Sd_o = d.Sd + (size_t)b * d.N;
  T* So_o = d.So + (size_t)b * d.N;
  T* Pd_o = d.Pd + (size_t)b * d.N;
  if (j == 0) {
    for (int i = 0; i < NX; ++i)
      for (int c = 0; c < NX; ++c) {
        Sd[i * NX + c] = -d.Gh[(size_t)(i * NM + c) * K + t];
        So_o[(size_t)(i * NX + c) * K] = T(0);
      }
    for (int i = 0; i < NX; ++i) gam[i] = (d.x[(size_t)i * K + t] - d.xs[(size_t)i * d.B + b]) - d.Gg[(size_t)i * K + t];
  } else {
    const size_t tp = t - 1;
    T AB[NX * NM];
    load_AB(d, tp, AB);
    // W = Ghat_{j-1} AB^T  (NM x NX)
    T W[NM * NX];
    for (int r = 0; r < NM; ++r) {
      T grow[NM];
      for (int c = 0; c < NM; ++c) grow[c] = d.Gh[(size_t)(r * NM + c) * K + tp];
      for (int i = 0; i < NX; ++i) {
        T acc = T(0);
        for (int c = 0; c < NM; ++c) acc += grow[c] * AB[i * NM + c];
        W[r * NX + i] = acc;
      }
    }
    for (int i = 0; i < NX; ++i)
      for (int c = 0; c < NX; ++c) {
        T acc = T(0);
        for (int r = 0; r < NM; ++r) acc += AB[i * NM + r] * W[r * NX + c];
        Sd[i * NX + c] = -(acc + d.Gh[(size_t)(i * NM + c) * K + t]);
      }
    // S_{j,j-1} = AB Ghat_{j-1}[:, :NX]
    for (int c = 0; c < NX; ++c) {
      T gcol[NM];
      for (int r = 0; r < NM; ++r) gcol[r] = d.Gh[(size_t)(r * NM + c) * K + tp];
      for (int i = 0; i < NX; ++i) {
        T acc = T(0);
        for (int r = 0; r < NM; ++r) acc += AB[i * NM + r] * gcol[r];
        So_o[(size_t)(i * NX + c) * K + j] = acc;
      }
    }
    T Ggp[NM];
    for (int r = 0; r < NM; ++r) Ggp[r] = d.Gg[(size_t)r * K + tp];
    for (int i = 0; i < NX; ++i) {
      T acc = T(0);
      for (int r = 0; r < NM; ++r) acc += AB[i * NM + r] * Ggp[r];
      T ck = d.x[(size_t)i * K + t] - d.xkp1[(size_t)i * K + tp];
      gam[i] = (ck + acc) - d.Gg[(size_t)i * K + t];
    }
  }
  for (int i = 0; i < NX; ++i) {
    d.gam[(size_t)i * K + t] = gam[i];
    for (int c = 0; c < NX; ++c) Sd_o[(size_t)(i * NX + c) * K + j] = Sd[i * NX + c];
  }
  if (jacobi) {
    for (int i = 0; i < NX; ++i)
      for (int c = 0; c < NX; ++c) Pd_o[(size_t)(i * NX + c) * K + j] = (i == c) ? T(1) / Sd[i * NX + i] : T(0);
  } else {
    for (int i = 0; i < NX * NX; ++i) Sd[i] = -Sd[i];     // -S_jj is SPD
    spd_inverse_inplace(Sd, NX, NX);
    for (int i = 0; i < NX; ++i)
      for (int c = 0; c < NX; ++c) Pd_o[(size_t)(i * NX + c) * K + j] = -Sd[i * NX + c];
  }
}

// -----------------------------------------------------------------------------------------------------------------
// Structured fast path (diag_mode): QuadraticCost with diagonal Q, QF, R.  G_k + rho I = diag(d) + gck gck^T, so
//   Ghat_k = diag(1/d) - s h h^T,  h = gck / d,  s = 1 / (1 + gck^T h)           (Sherman-Morrison)
// is stored as 2m+1 scalars per knot (rows [0,m) dinv, [m,2m) h, 2m: s of the Gh array) instead of m*m, and the Schur
// blocks are assembled from the integrator structure  AB = [[E0 + tau Ab], [Ab]],  Ab = [dt Dq, I + dt Dqd, dt Minv],
// E0 = [I, dte I, 0]  (euler: dte = dt, tau = 0; semi-implicit: dte = 0, tau = dt):
//   AB D AB^T = [[Z + tau (F + F^T) + tau^2 M, F^T + tau M], [F + tau M, M]],  M = Ab D Ab^T, F = Ab D E0^T, Z = E0 D E0^T.
// Same mathematics as k_kkt / k_schur / k_recover (formKKTSystemBlocks :216-260, solveKKTSystem_Schur :419-452), ~10x fewer flops.
// -----------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(128) k_kkt_diag(Dev<T> d, const int* list, const int* count) {
  const size_t gt = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int slot = (int)(gt / d.N);
  if (slot >= *count) return;
  const int k = (int)(gt % d.N);
  const int b = list[slot];
  const size_t t = (size_t)b * d.N + k;
  const size_t K = d.K;
  const bool terminal = (k == d.N - 1);
  T z[NM];
  load_xu(d.x, d.u, K, t, terminal, z, z + NX);
  const T* Q = cost_Q(d.cost, k, terminal);
  const T rho = d.rho[b];
  T g[NM], dd[NM];
  for (int i = 0; i < NX; ++i) {
    const T qi = Q[i * NX + i];
    g[i] = (z[i] - d.xg[(size_t)i * d.B + b]) * qi;
    dd[i] = qi + rho;
  }
  for (int i = 0; i < NU; ++i) {
    const T ri = d.cost.R[i * NU + i];
    g[NX + i] = terminal ? T(0) : z[NX + i] * ri;
    dd[NX + i] = ri + rho;
  }
  T gck[NM];
  for (int i = 0; i < NM; ++i) gck[i] = T(0);
  if (d.lim.any) {
    soft_grad(d.lim, z, d.mu + t, d.lam + t, K, terminal, gck);
    for (int i = 0; i < NM; ++i) g[i] += gck[i];
  }
  const int M = terminal ? NX : NM;
  T hg = T(0), den = T(1);
  T h[NM], dinv[NM];
  for (int i = 0; i < NM; ++i) {
    dinv[i] = (i < M) ? T(1) / dd[i] : T(0);
    h[i] = gck[i] * dinv[i];
    den += gck[i] * h[i];
    hg += h[i] * g[i];
  }
  const T sS = T(1) / den;
  for (int i = 0; i < NM; ++i) {
    d.Gh[(size_t)i * K + t] = dinv[i];
    d.Gh[(size_t)(NM + i) * K + t] = h[i];
    d.g[(size_t)i * K + t] = (i < M) ? g[i] : T(0);
    d.Gg[(size_t)i * K + t] = (i < M) ? dinv[i] * g[i] - sS * h[i] * hg : T(0);
  }
  d.Gh[(size_t)(2 * NM) * K + t] = sS;
}

// bottom rows of [A B]: Ab (NJ x NM row-major) = dt * dqdd + [0 I 0]   (both integrators)
template <typename T>
__device__ __forceinline__ void load_Ab(const Dev<T>& d, size_t t, T* Ab) {
  for (int a = 0; a < NJ; ++a) {
    for (int c = 0; c < NM; ++c) Ab[a * NM + c] = d.dt * d.dyn[(size_t)(a * 3 * NJ + c) * d.K + t];
    Ab[a * NM + NJ + a] += T(1);
  }
}

// k_schur_diag: one thread per block row j.  The knot's Ab (NJ x NM) is staged in shared memory ([entry][thread], conflict-free);
// M, F, v, w are accumulated in registers by streaming over the columns r of Ab (rank-1 updates M += d_r a_r a_r^T), and every
// entry of S_jj / S_j,j-1 / gamma_j is written straight to global memory -- no per-thread local arrays.
enum { SCHUR_THREADS = 64 };
template <typename T>
__global__ void __launch_bounds__(SCHUR_THREADS) k_schur_diag(Dev<T> d, const int* list, const int* count, int need_so) {
  extern __shared__ unsigned char smem_raw[];
  T* sAb = reinterpret_cast<T*>(smem_raw);            // [NJ*NM][SCHUR_THREADS]
  const int tx = threadIdx.x;
  const size_t gt = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int slot = (int)(gt / d.N);
  if (slot >= *count) return;
  const int j = (int)(gt % d.N);
  const int b = list[slot];
  const size_t t = (size_t)b * d.N + j;
  const size_t K = d.K;
  (void)0;
  T* Sd_o = d.Sd + (size_t)b * d.N;
  T* So_o = d.So + (size_t)b * d.N;
  T* gm_o = d.gam + (size_t)b * d.N;
  auto GH = [&](int e, size_t tt) -> T { return d.Gh[(size_t)e * K + tt]; };
  const T sj = GH(2 * NM, t);
  if (j == 0) {
    static_for<0, NX>([&](auto ic) {
      constexpr int i = decltype(ic)::value;
      const T hi = GH(NM + i, t);
      static_for<0, NX>([&](auto cc) {
        constexpr int c = decltype(cc)::value;
        const T val = -(((i == c) ? GH(i, t) : T(0)) - sj * hi * GH(NM + c, t));
        Sd_o[(size_t)(i * NX + c) * K + j] = val;
        So_o[(size_t)(i * NX + c) * K + j] = T(0);
      });
      gm_o[(size_t)i * K + j] = (d.x[(size_t)i * K + t] - d.xs[(size_t)i * d.B + b]) - d.Gg[(size_t)i * K + t];
    });
    return;
  }
  const size_t tp = t - 1;
  const T dte = d.integrator == 0 ? d.dt : T(0);
  const T tau = d.integrator == 0 ? T(0) : d.dt;
  const T sp = GH(2 * NM, tp);
  // stage Ab = dt * dqdd + [0 I 0] of knot j-1
  static_for<0, NJ>([&](auto ac) {
    constexpr int a = decltype(ac)::value;
    static_for<0, NM>([&](auto cc) {
      constexpr int c = decltype(cc)::value;
      sAb[(a * NM + c) * SCHUR_THREADS + tx] = d.dt * d.dyn[(size_t)(a * 3 * NJ + c) * K + tp] + ((c == NJ + a) ? T(1) : T(0));
    });
  });
  auto AB = [&](int a, int c) -> T { return sAb[(a * NM + c) * SCHUR_THREADS + tx]; };
  // stream over the columns r of Ab
  T Mp[NJ * (NJ + 1) / 2], F[NJ][NJ], vb[NJ], wb[NJ];      // Mp: packed lower triangle of M
  auto MI = [](int a, int c) { return a >= c ? a * (a + 1) / 2 + c : c * (c + 1) / 2 + a; };
  static_for<0, NJ>([&](auto ac) {
    constexpr int a = decltype(ac)::value;
    vb[a] = T(0); wb[a] = T(0);
    static_for<0, NJ>([&](auto cc) { constexpr int c = decltype(cc)::value; F[a][c] = T(0); });
    static_for<0, a + 1>([&](auto cc) { constexpr int c = decltype(cc)::value; Mp[a * (a + 1) / 2 + c] = T(0); });
  });
  static_for<0, NM>([&](auto rc) {
    constexpr int r = decltype(rc)::value;
    const T dr = GH(r, tp), hr = GH(NM + r, tp), gr = d.Gg[(size_t)r * K + tp];
    T col[NJ];
    static_for<0, NJ>([&](auto ac) { constexpr int a = decltype(ac)::value; col[a] = AB(a, r); });
    static_for<0, NJ>([&](auto ac) {
      constexpr int a = decltype(ac)::value;
      vb[a] += col[a] * hr;
      wb[a] += col[a] * gr;
      const T cd = col[a] * dr;
      static_for<0, a + 1>([&](auto cc) { constexpr int c = decltype(cc)::value; Mp[a * (a + 1) / 2 + c] += cd * col[c]; });
      if constexpr (r < NJ) F[a][r] += cd;                       // Ab[a][i] d_i
      else if constexpr (r < NX) F[a][r - NJ] += dte * cd;       // dte Ab[a][n+i] d_{n+i}
    });
  });
  // gamma_j = c_j + AB (Ghat g)_{j-1} - (Ghat g)_j[:nx]
  static_for<0, NJ>([&](auto ic) {
    constexpr int i = decltype(ic)::value;
    const T wt = d.Gg[(size_t)i * K + tp] + dte * d.Gg[(size_t)(NJ + i) * K + tp] + tau * wb[i];
    const T ckt = d.x[(size_t)i * K + t] - d.xkp1[(size_t)i * K + tp];
    gm_o[(size_t)i * K + j] = (ckt + wt) - d.Gg[(size_t)i * K + t];
    const T ckb = d.x[(size_t)(NJ + i) * K + t] - d.xkp1[(size_t)(NJ + i) * K + tp];
    gm_o[(size_t)(NJ + i) * K + j] = (ckb + wb[i]) - d.Gg[(size_t)(NJ + i) * K + t];
  });
  T v[NX];
  static_for<0, NJ>([&](auto ic) {
    constexpr int i = decltype(ic)::value;
    v[i] = GH(NM + i, tp) + dte * GH(NM + NJ + i, tp) + tau * vb[i];
    v[NJ + i] = vb[i];
  });
  // S_jj = -(AB D AB^T - sp v v^T + Ghat_j[:nx,:nx])
  static_for<0, NJ>([&](auto ic) {
    constexpr int i = decltype(ic)::value;
    const T Zi = GH(i, tp) + dte * dte * GH(NJ + i, tp);
    const T hxi = GH(NM + i, t), hxni = GH(NM + NJ + i, t);
    static_for<0, NJ>([&](auto cc) {
      constexpr int c = decltype(cc)::value;
      const T hxc = GH(NM + c, t), hxnc = GH(NM + NJ + c, t);
      const T mic = Mp[MI(i, c)];
      const T tt = ((i == c) ? Zi : T(0)) + tau * (F[c][i] + F[i][c]) + tau * tau * mic;
      const T tb = F[c][i] + tau * mic;                           // (top i, bottom c)
      const T gtt = ((i == c) ? GH(i, t) : T(0)) - sj * hxi * hxc;
      const T gtb = -sj * hxi * hxnc;
      const T gbb = ((i == c) ? GH(NJ + i, t) : T(0)) - sj * hxni * hxnc;
      Sd_o[(size_t)(i * NX + c) * K + j] = -((tt - sp * v[i] * v[c]) + gtt);
      Sd_o[(size_t)(i * NX + NJ + c) * K + j] = -((tb - sp * v[i] * v[NJ + c]) + gtb);
      Sd_o[(size_t)((NJ + c) * NX + i) * K + j] = -((tb - sp * v[NJ + c] * v[i]) + gtb);
      Sd_o[(size_t)((NJ + i) * NX + NJ + c) * K + j] = -((mic - sp * v[NJ + i] * v[NJ + c]) + gbb);
    });
  });
  // S_{j,j-1} = AB[:, :nx] diag(d_x) - sp v h_x^T   (not needed by the matrix-free PCG kernels)
  if (need_so) static_for<0, NX>([&](auto cc) {
    constexpr int c = decltype(cc)::value;
    const T dc = GH(c, tp), hc = GH(NM + c, tp);
    static_for<0, NJ>([&](auto ic) {
      constexpr int i = decltype(ic)::value;
      const T ab = AB(i, c);
      const T e0 = ((c == i) ? T(1) : T(0)) + ((c == NJ + i) ? dte : T(0));
      So_o[(size_t)(i * NX + c) * K + j] = (e0 + tau * ab) * dc - sp * v[i] * hc;
      So_o[(size_t)((NJ + i) * NX + c) * K + j] = ab * dc - sp * v[NJ + i] * hc;
    });
  });
}

// k_schur_rows: the same block row, NJ threads per knot -- thread (i, knot) stages row i of Ab and produces rows i and
// NJ+i of S_jj / S_j,j-1 / gamma_j (the body of k_schur_diag's outer loop over i).  k_schur_diag runs one 1700-FMA chain per thread at
// 8 warps per SM (255 registers, 55 KB of shared memory per 64 knots) and is pure latency (ncu: FP64 pipe 4 %, long-scoreboard stall
// 18.8 cycles per issue); here the chain is NJ times shorter, the thread needs ~1/4 of the registers, and the per-knot vectors
// (dinv, h, Ghat g of knot j-1, h of knot j) are staged once per knot instead of being re-read by every use.  Thread layout
// tx = i * SCHUR_KB + knot, so a warp is 32 consecutive knots of one row index: global accesses stay coalesced and the per-knot
// shared-memory records (odd stride) are conflict-free.  Every sum keeps k_schur_diag's operand order: bit-identical outputs.
enum { SCHUR_KB = 32 };
constexpr int SCHUR_REC = (NJ * NM + 3 * NM + NX + NJ * NJ + NX) | 1;   // Ab, dinv/h/Gg of knot j-1, h_x of knot j, F, v
template <typename T, int MINB>      // MINB: CTAs per SM the register allocation aims for (3: 96 registers, spills to L2; 2: 168, none -- measured 22.9 vs 31.5 ms)
__global__ void __launch_bounds__(SCHUR_KB * NJ, MINB) k_schur_rows(Dev<T> d, const int* list, const int* count, int need_so) {
  extern __shared__ unsigned char smem_raw[];
  const int i = threadIdx.x / SCHUR_KB, jj = threadIdx.x % SCHUR_KB;
  T* rec = reinterpret_cast<T*>(smem_raw) + (size_t)jj * SCHUR_REC;
  T* sAb = rec;                         // [NJ][NM]
  T* sD = sAb + NJ * NM;                // dinv of knot j-1
  T* sH = sD + NM;                      // h of knot j-1
  T* sG = sH + NM;                      // Ghat g of knot j-1
  T* sHx = sG + NM;                     // h[:nx] of knot j
  T* sF = sHx + NX;                     // [NJ][NJ]
  T* sv = sF + NJ * NJ;                 // [NX]
  const size_t gt = (size_t)blockIdx.x * SCHUR_KB + jj;
  const int slot = (int)(gt / d.N);
  const bool act = slot < *count;
  const int j = (int)(gt % d.N);
  const int b = act ? list[slot] : 0;
  const size_t t = (size_t)b * d.N + j;
  const size_t K = d.K;
  T* Sd_o = d.Sd + (size_t)b * d.N;
  T* So_o = d.So + (size_t)b * d.N;
  T* gm_o = d.gam + (size_t)b * d.N;
  auto GH = [&](int e, size_t tt) -> T { return d.Gh[(size_t)e * K + tt]; };
  const bool first = (j == 0);
  const bool work = act && !first;
  const size_t tp = work ? t - 1 : t;
  const T dte = d.integrator == 0 ? d.dt : T(0);
  const T tau = d.integrator == 0 ? T(0) : d.dt;
  if (act) {
    for (int e = i; e < NX; e += NJ) sHx[e] = GH(NM + e, t);
    if (work) {
      for (int e = i; e < NM; e += NJ) { sD[e] = GH(e, tp); sH[e] = GH(NM + e, tp); sG[e] = d.Gg[(size_t)e * K + tp]; }
      static_for<0, NM>([&](auto cc) {
        constexpr int c = decltype(cc)::value;
        sAb[i * NM + c] = d.dt * d.dyn[(size_t)(i * 3 * NJ + c) * K + tp] + ((c == NJ + i) ? T(1) : T(0));
      });
    }
  }
  __syncthreads();
  const T sj = act ? GH(2 * NM, t) : T(0);
  const T sp = work ? GH(2 * NM, tp) : T(0);
  T Mi[NJ], Fi[NJ], vbi = T(0), wbi = T(0);
  static_for<0, NJ>([&](auto cc) { constexpr int c = decltype(cc)::value; Mi[c] = T(0); Fi[c] = T(0); });
  T vi = T(0), vni = T(0);
  if (work) {
    static_for<0, NM>([&](auto rc) {
      constexpr int r = decltype(rc)::value;
      const T dr = sD[r], hr = sH[r], gr = sG[r];
      const T own = sAb[i * NM + r];      // own row re-read from the record instead of a private register copy
      vbi += own * hr;
      wbi += own * gr;
      const T cdi = own * dr;
      static_for<0, NJ>([&](auto cc) {
        constexpr int c = decltype(cc)::value;
        const T oc = sAb[c * NM + r];
        // entry (i, c) of the symmetric M: k_schur_diag accumulates (col[a] d_r) col[c] for a >= c
        const bool up = c > i;
        const T hi = up ? oc * dr : cdi, lo = up ? own : oc;
        Mi[c] += hi * lo;
      });
      if constexpr (r < NJ) Fi[r] += cdi;
      else if constexpr (r < NX) Fi[r - NJ] += dte * cdi;
    });
    // gamma_j = c_j + AB (Ghat g)_{j-1} - (Ghat g)_j[:nx]
    {
      const T wt = sG[i] + dte * sG[NJ + i] + tau * wbi;
      const T ckt = d.x[(size_t)i * K + t] - d.xkp1[(size_t)i * K + tp];
      gm_o[(size_t)i * K + j] = (ckt + wt) - d.Gg[(size_t)i * K + t];
      const T ckb = d.x[(size_t)(NJ + i) * K + t] - d.xkp1[(size_t)(NJ + i) * K + tp];
      gm_o[(size_t)(NJ + i) * K + j] = (ckb + wbi) - d.Gg[(size_t)(NJ + i) * K + t];
    }
    vi = sH[i] + dte * sH[NJ + i] + tau * vbi;
    vni = vbi;
    sv[i] = vi; sv[NJ + i] = vni;
    static_for<0, NJ>([&](auto cc) { constexpr int c = decltype(cc)::value; sF[i * NJ + c] = Fi[c]; });
  }
  __syncthreads();
  if (!act) return;
  if (first) {
    // S_00 = -Ghat_0[:nx,:nx], gamma_0 = (x_0 - xs) - (Ghat g)_0[:nx]; rows i and NJ+i
    for (int h = 0; h < 2; ++h) {
      const int ri = h * NJ + i;
      const T hi = sHx[ri], di = GH(ri, t);
      static_for<0, NX>([&](auto cc) {
        constexpr int c = decltype(cc)::value;
        const T val = -(((ri == c) ? di : T(0)) - sj * hi * sHx[c]);
        Sd_o[(size_t)(ri * NX + c) * K + j] = val;
        So_o[(size_t)(ri * NX + c) * K + j] = T(0);
      });
      gm_o[(size_t)ri * K + j] = (d.x[(size_t)ri * K + t] - d.xs[(size_t)ri * d.B + b]) - d.Gg[(size_t)ri * K + t];
    }
    return;
  }
  // S_jj = -(AB D AB^T - sp v v^T + Ghat_j[:nx,:nx])
  {
    const T Zi = sD[i] + dte * dte * sD[NJ + i];
    const T hxi = sHx[i], hxni = sHx[NJ + i];
    const T gdi = GH(i, t), gdni = GH(NJ + i, t);
    static_for<0, NJ>([&](auto cc) {
      constexpr int c = decltype(cc)::value;
      const T hxc = sHx[c], hxnc = sHx[NJ + c];
      const T mic = Mi[c];
      const T Fci = sF[c * NJ + i], Fic = Fi[c];
      const T vc = sv[c], vnc = sv[NJ + c];
      const T tt = ((i == c) ? Zi : T(0)) + tau * (Fci + Fic) + tau * tau * mic;
      const T tb = Fci + tau * mic;                               // (top i, bottom c)
      const T gtt = ((i == c) ? gdi : T(0)) - sj * hxi * hxc;
      const T gtb = -sj * hxi * hxnc;
      const T gbb = ((i == c) ? gdni : T(0)) - sj * hxni * hxnc;
      Sd_o[(size_t)(i * NX + c) * K + j] = -((tt - sp * vi * vc) + gtt);
      Sd_o[(size_t)(i * NX + NJ + c) * K + j] = -((tb - sp * vi * vnc) + gtb);
      Sd_o[(size_t)((NJ + c) * NX + i) * K + j] = -((tb - sp * vnc * vi) + gtb);
      Sd_o[(size_t)((NJ + i) * NX + NJ + c) * K + j] = -((mic - sp * vni * vnc) + gbb);
    });
  }
  // S_{j,j-1} = AB[:, :nx] diag(d_x) - sp v h_x^T   (not needed by the matrix-free PCG kernels)
  if (need_so) static_for<0, NX>([&](auto cc) {
    constexpr int c = decltype(cc)::value;
    const T dc = sD[c], hc = sH[c];
    const T ab = sAb[i * NM + c];
    const T e0 = ((c == i) ? T(1) : T(0)) + ((c == NJ + i) ? dte : T(0));
    So_o[(size_t)(i * NX + c) * K + j] = (e0 + tau * ab) * dc - sp * vi * hc;
    So_o[(size_t)((NJ + i) * NX + c) * K + j] = ab * dc - sp * vni * hc;
  });
}

// Pd_j = S_jj^-1 (block Jacobi / symmetric stair) or diag(S_jj)^-1 (Jacobi) -- PCG.compute_preconditioner (PCG.py:166-212).
// One thread per block row; -S_jj (SPD) is inverted in packed storage with a fully unrolled Cholesky, all in registers.
template <typename T>
__global__ void __launch_bounds__(128) k_pinv(Dev<T> d, const int* list, const int* count, int jacobi) {
  const size_t gt = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int slot = (int)(gt / d.N);
  if (slot >= *count) return;
  const int j = (int)(gt % d.N);
  const int b = list[slot];
  (void)0;
  const T* Sd_o = d.Sd + (size_t)b * d.N;
  T* Pd_o = d.Pd + (size_t)b * d.N;
  if (jacobi) {
    for (int i = 0; i < NX; ++i)
      for (int c = 0; c < NX; ++c) Pd_o[(size_t)(i * NX + c) * d.K + j] = (i == c) ? T(1) / Sd_o[(size_t)(i * NX + i) * d.K + j] : T(0);
    return;
  }
  T pk[NX * (NX + 1) / 2];
  static_for<0, NX>([&](auto ic) {
    constexpr int i = decltype(ic)::value;
    static_for<0, i + 1>([&](auto cc) { constexpr int c = decltype(cc)::value; pk[i * (i + 1) / 2 + c] = -Sd_o[(size_t)(i * NX + c) * d.K + j]; });
  });
  spd_inverse_packed<NX>(pk);
  static_for<0, NX>([&](auto ic) {
    constexpr int i = decltype(ic)::value;
    static_for<0, i + 1>([&](auto cc) {
      constexpr int c = decltype(cc)::value;
      const T val = -pk[i * (i + 1) / 2 + c];
      Pd_o[(size_t)(i * NX + c) * d.K + j] = val;
      Pd_o[(size_t)(c * NX + i) * d.K + j] = val;
    });
  });
}

// dz_k = Ghat_k (g_k - [l_k; 0] + AB_k^T l_{k+1}) of one knot (structured path)
template <typename T>
__device__ __forceinline__ void recover_diag_knot(const Dev<T>& d, int b, int k) {
  const size_t t = (size_t)b * d.N + k;
  const size_t K = d.K;
  const bool terminal = (k == d.N - 1);
  T rhs[NM];
  for (int i = 0; i < NM; ++i) rhs[i] = d.g[(size_t)i * K + t];
  const T* l = d.l + (size_t)b * d.N;
  for (int i = 0; i < NX; ++i) rhs[i] -= l[(size_t)i * K + k];
  if (!terminal) {
    const T dte = d.integrator == 0 ? d.dt : T(0);
    const T tau = d.integrator == 0 ? T(0) : d.dt;
    T ln[NX];
    for (int i = 0; i < NX; ++i) ln[i] = l[(size_t)i * K + k + 1];
    T wv[NJ];
    for (int a = 0; a < NJ; ++a) wv[a] = tau * ln[a] + ln[NJ + a];
    for (int c = 0; c < NM; ++c) {
      T acc = T(0);
      for (int a = 0; a < NJ; ++a) acc += (d.dt * d.dyn[(size_t)(a * 3 * NJ + c) * K + t] + ((c == NJ + a) ? T(1) : T(0))) * wv[a];
      rhs[c] += acc;
    }
    for (int i = 0; i < NJ; ++i) { rhs[i] += ln[i]; rhs[NJ + i] += dte * ln[i]; }
  }
  T hr = T(0);
  T dinv[NM], h[NM];
  for (int i = 0; i < NM; ++i) { dinv[i] = d.Gh[(size_t)i * K + t]; h[i] = d.Gh[(size_t)(NM + i) * K + t]; hr += h[i] * rhs[i]; }
  const T sS = d.Gh[(size_t)(2 * NM) * K + t];
  for (int i = 0; i < NM; ++i) d.dz[(size_t)i * K + t] = dinv[i] * rhs[i] - sS * h[i] * hr;
}

template <typename T>
__global__ void __launch_bounds__(128) k_recover_diag(Dev<T> d, const int* list, const int* count) {
  const size_t gt = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int slot = (int)(gt / d.N);
  if (slot >= *count) return;
  recover_diag_knot(d, list[slot], (int)(gt % d.N));
}

// dense reconstruction of Ghat from its structured form (B2T_ARR_GHAT in diag_mode)
template <typename T>
__global__ void k_fetch_ghat_diag(Dev<T> d, double* out) {
  const size_t gt = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gt >= d.K) return;
  const T sS = d.Gh[(size_t)(2 * NM) * d.K + gt];
  for (int i = 0; i < NM; ++i)
    for (int c = 0; c < NM; ++c) {
      const T v = ((i == c) ? d.Gh[(size_t)i * d.K + gt] : T(0)) - sS * d.Gh[(size_t)(NM + i) * d.K + gt] * d.Gh[(size_t)(NM + c) * d.K + gt];
      out[gt * NM * NM + i * NM + c] = (double)v;
    }
}

// -----------------------------------------------------------------------------------------------------------------
// k_pcg: one thread block per instance, one thread per row of the block-tridiagonal system (PCG.pcg, PCG.py:66-111).
// Preconditioners (PCG.py:166-212): J / BJ use the diagonal blocks; SS (symmetric stair) is applied in factored form
//   Pinv r = y - D^-1 (O y),  y = D^-1 r,  O = off-diagonal part of S     (== the reference's explicit
//   Pinv_{k,k-1} = -S_kk^-1 S_{k,k-1} S_{k-1,k-1}^-1).
// Dot products: warp-shuffle tree + fixed-order cross-warp sum (deterministic).
// -----------------------------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ T block_sum(T v, T* red, int tid, int nthreads) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = tid >> 5, nw = (nthreads + 31) >> 5;
  __syncthreads();                    // red[] free to overwrite
  if ((tid & 31) == 0) red[w] = v;
  __syncthreads();
  T s = T(0);
  for (int i = 0; i < nw; ++i) s += red[i];
  return s;
}

template <typename T>
__device__ __forceinline__ T bt_diag_row(const T* Md, const T* vec, int r, int j, size_t K) {
  T acc = T(0);
  const int i = r % NX;
#pragma unroll
  for (int c = 0; c < NX; ++c) acc += Md[(size_t)(i * NX + c) * K + j] * vec[j * NX + c];
  return acc;
}
template <typename T>
__device__ __forceinline__ T bt_off_row(const T* Mo, const T* vec, int r, int j, int i, size_t K, int N) {
  T acc = T(0);
  if (j > 0) {
#pragma unroll
    for (int c = 0; c < NX; ++c) acc += Mo[(size_t)(i * NX + c) * K + j] * vec[(j - 1) * NX + c];
  }
  if (j < N - 1) {
#pragma unroll
    for (int c = 0; c < NX; ++c) acc += Mo[(size_t)(c * NX + i) * K + j + 1] * vec[(j + 1) * NX + c];
  }
  return acc;
}

enum { PCG_MAX_RPT = 4 };   // rows per thread (R <= 4096)

template <typename T>
__global__ void __launch_bounds__(1024) k_pcg(Dev<T> d, const int* list, const int* count, int stair, T tol, int max_iter) {
  if ((int)blockIdx.x >= *count) return;
  const int b = list[blockIdx.x];
  const int N = d.N, R = N * NX;
  const int tid = threadIdx.x, nt = blockDim.x;
  extern __shared__ unsigned char smem_raw[];
  T* p_s = reinterpret_cast<T*>(smem_raw);
  T* y_s = p_s + R;
  T* w_s = y_s + R;
  T* red = w_s + R;
  const T* Sd = d.Sd + (size_t)b * d.N;
  const T* So = d.So + (size_t)b * d.N;
  const T* Pd = d.Pd + (size_t)b * d.N;
  const T* gam = d.gam + (size_t)b * d.N;
  T rr[PCG_MAX_RPT], xx[PCG_MAX_RPT], pp[PCG_MAX_RPT], rt[PCG_MAX_RPT];
#pragma unroll
  for (int m = 0; m < PCG_MAX_RPT; ++m) { rr[m] = T(0); xx[m] = T(0); pp[m] = T(0); rt[m] = T(0); }

  auto precond = [&]() {
    // input r in rr[], output rt[]; uses y_s, w_s
    for (int m = 0, r = tid; r < R; r += nt, ++m) y_s[r] = rr[m];
    __syncthreads();
    T yv[PCG_MAX_RPT];
    for (int m = 0, r = tid; r < R; r += nt, ++m) yv[m] = bt_diag_row(Pd, y_s, r, r / NX, d.K);
    if (!stair) {
      for (int m = 0, r = tid; r < R; r += nt, ++m) rt[m] = yv[m];
      __syncthreads();
      return;
    }
    __syncthreads();
    for (int m = 0, r = tid; r < R; r += nt, ++m) w_s[r] = yv[m];      // w_s = y
    __syncthreads();
    T ov[PCG_MAX_RPT];
    for (int m = 0, r = tid; r < R; r += nt, ++m) ov[m] = bt_off_row(So, w_s, r, r / NX, r % NX, d.K, N);
    __syncthreads();
    for (int m = 0, r = tid; r < R; r += nt, ++m) y_s[r] = ov[m];      // y_s = O y
    __syncthreads();
    for (int m = 0, r = tid; r < R; r += nt, ++m) rt[m] = yv[m] - bt_diag_row(Pd, y_s, r, r / NX, d.K);
    __syncthreads();
  };

  // x0 = 0  ->  r = b
  for (int m = 0, r = tid; r < R; r += nt, ++m) { rr[m] = gam[(size_t)(r % NX) * d.K + r / NX]; xx[m] = T(0); }
  precond();
  T part = T(0);
  for (int m = 0, r = tid; r < R; r += nt, ++m) { pp[m] = rt[m]; part += rr[m] * rt[m]; }
  T nu = block_sum(part, red, tid, nt);
  if (tid == 0 && d.nu_trace) d.nu_trace[(size_t)b * NU_TRACE_LEN] = fabs(nu);
  int iters = 0;
  for (int it = 0; it < max_iter; ++it) {
    for (int m = 0, r = tid; r < R; r += nt, ++m) p_s[r] = pp[m];
    __syncthreads();
    T ap[PCG_MAX_RPT];
    part = T(0);
    for (int m = 0, r = tid; r < R; r += nt, ++m) {
      const int j = r / NX, i = r % NX;
      ap[m] = bt_diag_row(Sd, p_s, r, j, d.K) + bt_off_row(So, p_s, r, j, i, d.K, N);
      part += pp[m] * ap[m];
    }
    const T pAp = block_sum(part, red, tid, nt);
    const T alpha = nu / pAp;
    for (int m = 0, r = tid; r < R; r += nt, ++m) { rr[m] -= ap[m] * alpha; xx[m] += pp[m] * alpha; }
    precond();
    part = T(0);
    for (int m = 0, r = tid; r < R; r += nt, ++m) part += rr[m] * rt[m];
    const T nu_prime = block_sum(part, red, tid, nt);
    iters = it + 1;
    if (tid == 0 && d.nu_trace && iters < NU_TRACE_LEN) d.nu_trace[(size_t)b * NU_TRACE_LEN + iters] = fabs(nu_prime);
    if (fabs(nu_prime) < tol) break;
    const T beta = nu_prime / nu;
    for (int m = 0, r = tid; r < R; r += nt, ++m) pp[m] = rt[m] + pp[m] * beta;
    nu = nu_prime;
  }
  for (int m = 0, r = tid; r < R; r += nt, ++m) d.l[(size_t)(r % NX) * d.K + (size_t)b * N + r / NX] = xx[m];
  if (tid == 0) {
    d.pcg_iters[b] = iters;
    d.tot_pcg[b] += iters;
    d.tot_qp[b] += 1;
  }
}

// -----------------------------------------------------------------------------------------------------------------
// k_pcg2: register / shared-memory tiled PCG.  One block per instance, TB = NX/RPT threads per block row, RPT rows per
// thread.  The off-diagonal blocks of S (used four times per iteration: S p down/up, O y down/up) live in REGISTERS
// (2*RPT*NX values per thread); the diagonal blocks of S and of the preconditioner live in SHARED memory in a
// thread-major layout (conflict-free); vectors are exchanged through two zero-padded shared buffers, so a thread loads
// each vector entry once for its RPT rows.  Same arithmetic as k_pcg (PCG.py:66-111), same deterministic reductions.
// SMEM = false streams the diagonal blocks from global memory (L1/L2) when N*NX*NX*2 scalars exceed shared memory.
// -----------------------------------------------------------------------------------------------------------------
template <typename T, int RPT, int CS, bool SMEM, int MAXT>
__global__ void __launch_bounds__(MAXT) k_pcg2(Dev<T> d, const int* list, const int* count, int stair, T tol, int max_iter) {
  if ((int)blockIdx.x >= *count) return;
  // thread (j, g, h): block row j, rows i0 = g*RPT .. i0+RPT-1, columns c0 = h*NXC .. c0+NXC-1 of every 1 x NX row slice.
  // CS = 2 halves the registers per thread (more resident warps, room for the compiler to pipeline loads); the CS partial
  // sums of a row are combined with one butterfly shuffle.
  constexpr int TB = NX / RPT;                      // row groups per block row
  constexpr int TBC = TB * CS;                      // threads per block row (all in one warp)
  constexpr int NXC = NX / CS;                      // columns per thread
  static_assert(NX % RPT == 0 && NX % CS == 0 && NXC % 2 == 0, "tile shape");
  static_assert(32 % TBC == 0, "the threads of a block row must share a warp");
  static_assert(CS == 1 || CS == 2, "column split");
  const int b = list[blockIdx.x];
  const int N = d.N, NT = N * TBC;
  const size_t K = d.K;
  const int tid = threadIdx.x, nt = blockDim.x;
  const bool live = tid < NT;
  const int j = live ? tid / TBC : 0;
  const int i0 = live ? ((tid % TBC) / CS) * RPT : 0;
  const int h = tid % CS;
  const int c0 = h * NXC;
  const bool lead = live && h == 0;                  // the lane that stores / contributes to dot products
  extern __shared__ unsigned char smem_raw[];
  const int VL = (N + 2) * NX;                       // padded vector length: blocks 0 and N+1 stay zero
  T* A_s = reinterpret_cast<T*>(smem_raw);          // neighbour exchange (p, y)
  T* B_s = A_s + VL;                                 // own-block exchange (r)
  T* C_s = B_s + VL;                                 // own-block exchange (O y)
  T* red = C_s + VL;                                 // 32
  // diagonal blocks: [RPT][NXC/2][MAXT] pairs (columns c, c+1 of this thread's slice) -> one 128-bit load per pair, conflict-free
  using T2 = typename std::conditional<sizeof(T) == 8, double2, float2>::type;
  T2* Sd_s = reinterpret_cast<T2*>(red + 32);
  T2* Pd_s = Sd_s + (SMEM ? (size_t)RPT * (NXC / 2) * MAXT : 0);
  const T* Sd = d.Sd + (size_t)b * d.N;
  const T* So = d.So + (size_t)b * d.N;
  const T* Pd = d.Pd + (size_t)b * d.N;
  const T* gam = d.gam + (size_t)b * d.N;
  // ---- one-time loads: off-diagonal blocks into registers, diagonal blocks into shared memory
  T so_dn[RPT][NXC], so_up[RPT][NXC];
#pragma unroll
  for (int k = 0; k < RPT; ++k)
#pragma unroll
    for (int cc = 0; cc < NXC; ++cc) {
      const int c = c0 + cc;
      so_dn[k][cc] = live ? So[(size_t)((i0 + k) * NX + c) * K + j] : T(0);                                        // S_{j,j-1}[i][c]  (zero for j = 0)
      so_up[k][cc] = (live && j < N - 1) ? So[(size_t)(c * NX + i0 + k) * K + j + 1] : T(0);   // S_{j+1,j}[c][i]
      if constexpr (SMEM) {
        if (cc % 2 == 0) {
          T2 sv, pv;
          sv.x = live ? Sd[(size_t)((i0 + k) * NX + c) * K + j] : T(0); sv.y = live ? Sd[(size_t)((i0 + k) * NX + c + 1) * K + j] : T(0);
          pv.x = live ? Pd[(size_t)((i0 + k) * NX + c) * K + j] : T(0); pv.y = live ? Pd[(size_t)((i0 + k) * NX + c + 1) * K + j] : T(0);
          Sd_s[(k * (NXC / 2) + cc / 2) * MAXT + tid] = sv;
          Pd_s[(k * (NXC / 2) + cc / 2) * MAXT + tid] = pv;
        }
      }
    }
  for (int idx = tid; idx < 3 * VL; idx += nt) A_s[idx] = T(0);
  __syncthreads();
  // pair (cc, cc+1) of this thread's column slice, cc even
  auto ldSd2 = [&](int k, int cc) -> T2 {
    if constexpr (SMEM) return Sd_s[(k * (NXC / 2) + cc / 2) * MAXT + tid];
    else { T2 v; v.x = Sd[(size_t)((i0 + k) * NX + c0 + cc) * K + j]; v.y = Sd[(size_t)((i0 + k) * NX + c0 + cc + 1) * K + j]; return v; }
  };
  auto ldPd2 = [&](int k, int cc) -> T2 {
    if constexpr (SMEM) return Pd_s[(k * (NXC / 2) + cc / 2) * MAXT + tid];
    else { T2 v; v.x = Pd[(size_t)((i0 + k) * NX + c0 + cc) * K + j]; v.y = Pd[(size_t)((i0 + k) * NX + c0 + cc + 1) * K + j]; return v; }
  };
  auto ldV2 = [&](const T* V, int idx) -> T2 { return *reinterpret_cast<const T2*>(V + idx); };   // idx even -> 16-byte aligned
  auto combine = [&](T v) -> T {
    if constexpr (CS == 2) v += __shfl_xor_sync(0xffffffffu, v, 1);
    return v;
  };
  const int vb = (j + 1) * NX;                       // own block in the padded vectors
  // out = Pd_jj * V[own block]; the block's entries were written by lanes of this warp (-> __syncwarp suffices)
  auto pd_mul = [&](const T* V, T* out) {
    T o0[RPT], o1[RPT];
#pragma unroll
    for (int k = 0; k < RPT; ++k) { o0[k] = T(0); o1[k] = T(0); }
#pragma unroll
    for (int cc = 0; cc < NXC; cc += 2) {
      const T2 v = ldV2(V, vb + c0 + cc);
#pragma unroll
      for (int k = 0; k < RPT; ++k) { const T2 m2 = ldPd2(k, cc); o0[k] += m2.x * v.x; o1[k] += m2.y * v.y; }
    }
#pragma unroll
    for (int k = 0; k < RPT; ++k) out[k] = combine(o0[k] + o1[k]);
  };
  auto store = [&](T* V, const T* val) {
    if (lead) {
#pragma unroll
      for (int k = 0; k < RPT; ++k) V[vb + i0 + k] = val[k];
    }
  };
  T rr[RPT], xx[RPT], pp[RPT], rt[RPT], yv[RPT], tmp[RPT];
  // rt = Pinv rr.  Block barriers: one (stair only), before the neighbour reads of y.
  auto precond = [&]() {
    store(B_s, rr);
    __syncwarp();
    pd_mul(B_s, yv);
    if (!stair) {
#pragma unroll
      for (int k = 0; k < RPT; ++k) rt[k] = yv[k];
      return;
    }
    store(A_s, yv);
    __syncthreads();
    T t1[RPT], t2[RPT];
#pragma unroll
    for (int k = 0; k < RPT; ++k) { t1[k] = T(0); t2[k] = T(0); }
#pragma unroll
    for (int cc = 0; cc < NXC; cc += 2) {
      const T2 ym = ldV2(A_s, vb - NX + c0 + cc), yp = ldV2(A_s, vb + NX + c0 + cc);
#pragma unroll
      for (int k = 0; k < RPT; ++k) {
        t1[k] += so_dn[k][cc] * ym.x; t2[k] += so_up[k][cc] * yp.x;
        t1[k] += so_dn[k][cc + 1] * ym.y; t2[k] += so_up[k][cc + 1] * yp.y;
      }
    }
#pragma unroll
    for (int k = 0; k < RPT; ++k) tmp[k] = combine(t1[k] + t2[k]);
    store(C_s, tmp);
    __syncwarp();
    pd_mul(C_s, tmp);
#pragma unroll
    for (int k = 0; k < RPT; ++k) rt[k] = yv[k] - tmp[k];
  };
#pragma unroll
  for (int k = 0; k < RPT; ++k) { rr[k] = live ? gam[(size_t)(i0 + k) * K + j] : T(0); xx[k] = T(0); }
  precond();
  T part = T(0);
#pragma unroll
  for (int k = 0; k < RPT; ++k) { pp[k] = rt[k]; part += rr[k] * rt[k]; }
  T nu = block_sum(lead ? part : T(0), red, tid, nt);   // its barriers also order the y reads above before the p store below
  if (tid == 0 && d.nu_trace) d.nu_trace[(size_t)b * NU_TRACE_LEN] = fabs(nu);
  int iters = 0;
  for (int it = 0; it < max_iter; ++it) {
    store(A_s, pp);
    __syncthreads();
    T a0[RPT], a1[RPT], a2[RPT], ap[RPT];
#pragma unroll
    for (int k = 0; k < RPT; ++k) { a0[k] = T(0); a1[k] = T(0); a2[k] = T(0); }
#pragma unroll
    for (int cc = 0; cc < NXC; cc += 2) {
      const T2 pm = ldV2(A_s, vb - NX + c0 + cc), p0 = ldV2(A_s, vb + c0 + cc), pq = ldV2(A_s, vb + NX + c0 + cc);
#pragma unroll
      for (int k = 0; k < RPT; ++k) {
        const T2 m2 = ldSd2(k, cc);
        a0[k] += m2.x * p0.x; a1[k] += so_dn[k][cc] * pm.x; a2[k] += so_up[k][cc] * pq.x;
        a0[k] += m2.y * p0.y; a1[k] += so_dn[k][cc + 1] * pm.y; a2[k] += so_up[k][cc + 1] * pq.y;
      }
    }
    part = T(0);
#pragma unroll
    for (int k = 0; k < RPT; ++k) { ap[k] = combine(a0[k] + a1[k] + a2[k]); part += pp[k] * ap[k]; }
    const T pAp = block_sum(lead ? part : T(0), red, tid, nt);     // barriers: all reads of p from A_s are complete
    const T alpha = nu / pAp;
#pragma unroll
    for (int k = 0; k < RPT; ++k) { rr[k] -= ap[k] * alpha; xx[k] += pp[k] * alpha; }
    precond();
    part = T(0);
#pragma unroll
    for (int k = 0; k < RPT; ++k) part += rr[k] * rt[k];
    const T nu_prime = block_sum(lead ? part : T(0), red, tid, nt);
    iters = it + 1;
    if (tid == 0 && d.nu_trace && iters < NU_TRACE_LEN) d.nu_trace[(size_t)b * NU_TRACE_LEN + iters] = fabs(nu_prime);
    if (fabs(nu_prime) < tol) break;
    const T beta = nu_prime / nu;
#pragma unroll
    for (int k = 0; k < RPT; ++k) pp[k] = rt[k] + pp[k] * beta;
    nu = nu_prime;
  }
  if (lead) {
#pragma unroll
    for (int k = 0; k < RPT; ++k) d.l[(size_t)(i0 + k) * K + (size_t)b * N + j] = xx[k];
  }
  if (tid == 0) {
    d.pcg_iters[b] = iters;
    d.tot_pcg[b] += iters;
    d.tot_qp[b] += 1;
  }
}

// -----------------------------------------------------------------------------------------------------------------
// k_pcg3: matrix-free, register-resident PCG for the structured (diag_mode) path.  S and the stair off-diagonal blocks
// are never read: with  Ghat_k = diag(dinv) - s h h^T  and  AB_k = [[E0 + tau Ab],[Ab]]  (see k_kkt_diag)
//     w_k = Ghat_k ([p_k;0] - AB_k^T p_{k+1}),     (S p)_{k+1} = AB_k w_k - w_{k+1}[:nx],   (S p)_0 = -w_0[:nx]
//     q_k = Ghat_k [y_k;0],  q'_k = Ghat_k AB_k^T y_{k+1},   (O y)_{k+1} = AB_k q_k + q'_{k+1}[:nx],  (O y)_0 = q'_0[:nx]
// (O = off-diagonal part of S, used by the symmetric-stair preconditioner  Pinv r = y - D^-1 O y,  y = D^-1 r).
// Four lanes per knot k: lane g holds columns [g*MC, g*MC+MC) of Ab_k (all NJ rows), the matching entries of dinv_k, h_k,
// and rows [g*RPT, g*RPT+RPT) of the preconditioner block of block row j = (k+1) mod N, whose PCG vector entries it owns.
// Everything matrix-like stays in registers for all iterations; shared memory only carries the vectors between lanes.
// Same PCG recurrence and exit test as PCG.pcg (PCG.py:66-111); the products are algebraically identical to S p / O y
// (verified against the explicit form: identical iteration counts, SURVEY.md 7.2 parity floor).
// -----------------------------------------------------------------------------------------------------------------
constexpr int PCG3_NMS = (NM + 3) / 4 * 4;   // 18 -> 20 doubles: rows of consecutive knots start 4 double-banks apart
template <typename T, int MAXT, bool PDS, int LPK, bool PDC = false>
__global__ void __launch_bounds__(MAXT, PDS ? 2 : 1) k_pcg3(Dev<T> d, const int* list, const int* count, int stair, T tol, int max_iter) {
  if ((int)blockIdx.x >= *count) return;
  // LPK lanes per knot (2 or 4).  Only launched when nx % (2 LPK) == 0; the max() keeps the definition well-formed otherwise.
  constexpr int RPT = (NX % LPK == 0) ? NX / LPK : 1;   // owned rows per lane; divides NJ
  constexpr int MC = (NM + LPK - 1) / LPK;               // Ab columns per lane
  constexpr int NMS = PCG3_NMS;                          // padded row stride of W / Wq: bank-conflict-free 64-bit accesses
  using T2 = typename std::conditional<sizeof(T) == 8, double2, float2>::type;
  const int b = list[blockIdx.x];
  const int N = d.N;
  const size_t K = d.K;
  const int tid = threadIdx.x, nt = blockDim.x;
  const bool live = tid < LPK * N;
  const int k = live ? tid / LPK : 0;         // knot
  const int g = tid % LPK;                    // lane within the knot group
  const bool has_next = live && (k < N - 1);  // knot N-1 has no dynamics; its group owns block row 0
  const int jo = (k + 1 == N) ? 0 : k + 1;    // owned block row
  const int c0 = g * MC, i0 = g * RPT;
  const size_t tk = (size_t)b * N + k, tj = (size_t)b * N + jo;
  const T dte = d.integrator == 0 ? d.dt : T(0);
  const T tau = d.integrator == 0 ? T(0) : d.dt;
  extern __shared__ unsigned char smem_raw[];
  T* V = reinterpret_cast<T*>(smem_raw);      // [(N+1)][NX]  p / r / O y; block N stays zero
  T* V2 = V + (N + 1) * NX;                   // [(N+1)][NX]  y; block N stays zero
  T* W = V2 + (N + 1) * NX;                   // [N][NMS]     w or q'
  T* Wq = W + N * NMS;                        // [N][NMS]     q
  T* red = Wq + N * NMS;                      // 2 x 32: double-buffered warp partial sums (zero beyond the launched warps)
  // PDS: the preconditioner rows live in shared memory ([RPT][NX/2][MAXT] pairs, conflict-free 128-bit loads) instead of registers,
  // so that TWO instances are resident per SM
  T2* Pd_s = reinterpret_cast<T2*>(red + 64);
  // ---- resident data (registers for all iterations)
  T ab[NJ][MC], dinv[MC], hh[MC], pd[PDS ? 1 : RPT][PDS ? 2 : NX];
  // per-column constants, computed once: validity, the E0^T coupling (which entry of z_q feeds column c, with which factor) and
  // the self term ([p_k; 0])_c
  bool cval[MC], sval[MC];
  T emul[MC];
  int eidx[MC], sidx[MC];
#pragma unroll
  for (int i = 0; i < MC; ++i) {
    const int c = c0 + i;
    cval[i] = live && c < NM;
    dinv[i] = cval[i] ? d.Gh[(size_t)c * K + tk] : T(0);
    hh[i] = cval[i] ? d.Gh[(size_t)(NM + c) * K + tk] : T(0);
#pragma unroll
    for (int a = 0; a < NJ; ++a)
      ab[a][i] = (cval[i] && has_next) ? d.dt * d.dyn[(size_t)(a * 3 * NJ + c) * K + tk] + ((c == NJ + a) ? T(1) : T(0)) : T(0);
    emul[i] = (!has_next || !cval[i] || c >= NX) ? T(0) : (c < NJ ? T(1) : dte);
    eidx[i] = (c < NJ) ? c : ((c < NX) ? c - NJ : 0);
    sval[i] = cval[i] && c < NX;
    sidx[i] = (c < NX) ? c : 0;
  }
  const T sS = live ? d.Gh[(size_t)(2 * NM) * K + tk] : T(0);
  // without soft limits gck = 0: Ghat is diagonal (k_kkt_diag) and the h-reductions vanish.  With limits the same holds for every knot
  // without a violated bound (h_k = 0, so h^T u = 0 exactly): a warp whose 8 knots all have h = 0 skips the shuffles -- same values
  bool hnz = false;
#pragma unroll
  for (int i = 0; i < MC; ++i) hnz = hnz || (hh[i] != T(0));
  const bool rank1 = d.lim.any != 0 && __any_sync(0xffffffffu, hnz);
#pragma unroll
  for (int r = 0; r < RPT; ++r)
#pragma unroll
    for (int c = 0; c < NX; c += 2) {
      const T p0 = live ? d.Pd[(size_t)((i0 + r) * NX + c) * K + tj] : T(0);
      const T p1 = live ? d.Pd[(size_t)((i0 + r) * NX + c + 1) * K + tj] : T(0);
      if constexpr (PDS) { T2 v; v.x = p0; v.y = p1; Pd_s[(r * (NX / 2) + c / 2) * MAXT + tid] = v; }
      else { pd[r][c] = p0; pd[r][c + 1] = p1; }
    }
  for (int idx = tid; idx < 2 * (N + 1) * NX + 2 * N * NMS + 64; idx += nt) V[idx] = T(0);
  __syncthreads();
  const bool top = i0 < NJ;                   // rows i0.. are q rows (top half of [A B]) for the first LPK/2 lanes, qd rows for the others
  const bool odd = (i0 % NJ) != 0;            // row % NJ = (i0 % NJ) + r; with LPK = 4 that is RPT + r for the odd lanes, with LPK = 2 always r
  const int ownV = jo * NX, ownW = jo * NMS + i0, kV = k * NX, nV = (k + 1) * NX, kW = k * NMS;

  auto quad = [&](T v) -> T {                 // sum over the LPK lanes of the knot group (all lanes get it)
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    if constexpr (LPK == 4) v += __shfl_xor_sync(0xffffffffu, v, 2);
    return v;
  };
  // deterministic block sum: warp butterfly, one barrier, pairwise tree over the warp partials; `red` alternates between two buffers
  // (the barrier of the next reduction orders the reads of this one before the buffer is written again)
  int red_sel = 0;
  auto bsum = [&](T v) -> T {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    T* rb = red + 32 * red_sel;
    red_sel ^= 1;
    if ((tid & 31) == 0) rb[tid >> 5] = v;
    __syncthreads();
    constexpr int NW = MAXT / 32;
    T t[NW];
#pragma unroll
    for (int i = 0; i < NW; i += 2) { const T2 p2 = *reinterpret_cast<const T2*>(rb + i); t[i] = p2.x; t[i + 1] = p2.y; }
#pragma unroll
    for (int st = 1; st < NW; st *= 2)
#pragma unroll
      for (int i = 0; i + st < NW; i += 2 * st) t[i] += t[i + st];
    return t[0];
  };
  auto publish = [&](T* buf, const T* val) {  // owned rows of block jo
    if (live) {
#pragma unroll
      for (int r = 0; r < RPT; ++r) buf[ownV + i0 + r] = val[r];
    }
  };
  // tc_i = (AB_k^T z)_c for my columns, z = block k+1 of `buf` (the zero block for k = N-1)
  auto abt = [&](const T* buf, T* tc) {
    T pi[NJ];
#pragma unroll
    for (int a = 0; a < NJ; a += 2) {
      const T2 zq = *reinterpret_cast<const T2*>(buf + nV + a);
      const T2 zv = *reinterpret_cast<const T2*>(buf + nV + NJ + a);
      pi[a] = tau * zq.x + zv.x;
      if (a + 1 < NJ) pi[a + 1] = tau * zq.y + zv.y;
    }
#pragma unroll
    for (int i = 0; i < MC; ++i) {      // two accumulators per column: half the dependent-FMA depth
      T acc0 = emul[i] * buf[nV + eidx[i]], acc1 = T(0);
#pragma unroll
      for (int a = 0; a < NJ; a += 2) {
        acc0 += ab[a][i] * pi[a];
        if (a + 1 < NJ) acc1 += ab[a + 1][i] * pi[a + 1];
      }
      tc[i] = acc0 + acc1;
    }
  };
  // out_r (owned rows of block jo) = (AB_k zz)_row + sign * Wn[jo][row];  zz: my columns in zc (registers), all columns in `full`
  auto abmul = [&](const T* zc, const T* full, const T* Wn, T sign, T* out) {
    T pb[NJ];
#pragma unroll
    for (int a = 0; a < NJ; ++a) {
      T acc = T(0);
#pragma unroll
      for (int i = 0; i < MC; ++i) acc += ab[a][i] * zc[i];
      pb[a] = acc;
    }
    // the lanes of the group only need the RPT rows congruent to their own: with four lanes, exchange halves with the xor-1
    // partner (reduce-scatter), then all-reduce over xor-2 -- same summation order as a plain butterfly, half the shuffles
    T bot[RPT];
    if constexpr (LPK == 4) {
#pragma unroll
      for (int r = 0; r < RPT; ++r) {
        const T send = odd ? pb[r] : pb[RPT + r];
        T keep = odd ? pb[RPT + r] : pb[r];
        keep += __shfl_xor_sync(0xffffffffu, send, 1);
        keep += __shfl_xor_sync(0xffffffffu, keep, 2);
        bot[r] = keep;
      }
    } else {
#pragma unroll
      for (int r = 0; r < RPT; ++r) bot[r] = quad(pb[r % NJ]);
    }
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
      // top rows: z_q + dte z_qd + tau (Ab z);  bottom rows: Ab z.  full[kW + row], full[kW + NJ + row] only matter for top rows
      const T zq = full[kW + (top ? i0 + r : 0)], zd = full[kW + NJ + (top ? i0 + r : 0)];
      const T val = top ? (zq + dte * zd + tau * bot[r]) : bot[r];
      out[r] = has_next ? val + sign * Wn[ownW + r] : (live ? sign * Wn[ownW + r] : T(0));
    }
  };
  // out = Pd_jo * buf[jo]  (entries published by the lanes of this group: __syncwarp suffices)
  auto pd_mul = [&](const T* buf, T* out) {
    T o0[RPT], o1[RPT];                       // two accumulators per row (four were measured 2 % slower)
#pragma unroll
    for (int r = 0; r < RPT; ++r) { o0[r] = T(0); o1[r] = T(0); }
#pragma unroll
    for (int c = 0; c < NX; c += 2) {
      const T2 v = *reinterpret_cast<const T2*>(buf + ownV + c);
#pragma unroll
      for (int r = 0; r < RPT; ++r) {
        if constexpr (PDS) { const T2 m2 = Pd_s[(r * (NX / 2) + c / 2) * MAXT + tid]; o0[r] += m2.x * v.x; o1[r] += m2.y * v.y; }
        else { o0[r] += pd[r][c] * v.x; o1[r] += pd[r][c + 1] * v.y; }
      }
    }
#pragma unroll
    for (int r = 0; r < RPT; ++r) out[r] = o0[r] + o1[r];
  };
  // PDC (experiment, B2T_PCG_VARIANT=7): D^-1 is symmetric, so the lane's RPT rows are also its RPT COLUMNS: partial products from the
  // lane's own RPT entries of the vector (registers, no publish / load), then a reduce-scatter over the four lanes (9 shuffles)
  auto pd_mul_col = [&](const T* val, T* out) {
    T y[NX];
#pragma unroll
    for (int j = 0; j < NX; ++j) {
      T acc = T(0);
#pragma unroll
      for (int r = 0; r < RPT; ++r) acc += pd[r][j] * val[r];
      y[j] = acc;
    }
    constexpr int H = NX / 2;
    const bool hi = (g & 2) != 0, od = (g & 1) != 0;
    T keep[H];
#pragma unroll
    for (int j = 0; j < H; ++j) {
      const T send = hi ? y[j] : y[H + j];
      keep[j] = (hi ? y[H + j] : y[j]) + __shfl_xor_sync(0xffffffffu, send, 2);
    }
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
      const T send = od ? keep[r] : keep[RPT + r];
      out[r] = (od ? keep[RPT + r] : keep[r]) + __shfl_xor_sync(0xffffffffu, send, 1);
    }
  };
  T rr[RPT], xx[RPT], pp[RPT], rt[RPT], yv[RPT], tmp[RPT];
  auto precond = [&]() {
    if constexpr (PDC && !PDS && LPK == 4) {
      pd_mul_col(rr, yv);
    } else {
      publish(V, rr);
      __syncwarp();
      pd_mul(V, yv);
    }
    if (!stair) {
#pragma unroll
      for (int r = 0; r < RPT; ++r) rt[r] = yv[r];
      return;
    }
    publish(V2, yv);
    __syncthreads();
    // q_c = Ghat [y_k;0], q'_c = Ghat AB^T y_{k+1}
    T u1[MC], u2[MC], h1 = T(0), h2 = T(0);
    abt(V2, u2);
#pragma unroll
    for (int i = 0; i < MC; ++i) {
      u1[i] = sval[i] ? V2[kV + sidx[i]] : T(0);
      h1 += hh[i] * u1[i];
      h2 += hh[i] * u2[i];
    }
    if (rank1) { h1 = sS * quad(h1); h2 = sS * quad(h2); } else { h1 = T(0); h2 = T(0); }
#pragma unroll
    for (int i = 0; i < MC; ++i) {
      u1[i] = dinv[i] * u1[i] - hh[i] * h1;
      u2[i] = dinv[i] * u2[i] - hh[i] * h2;
      if (cval[i]) { Wq[kW + c0 + i] = u1[i]; W[kW + c0 + i] = u2[i]; }
    }
    __syncthreads();
    abmul(u1, Wq, W, T(1), tmp);
    if constexpr (PDC && !PDS && LPK == 4) {
      T t2[RPT];
      pd_mul_col(tmp, t2);
#pragma unroll
      for (int r = 0; r < RPT; ++r) tmp[r] = t2[r];
    } else {
      publish(V, tmp);
      __syncwarp();
      pd_mul(V, tmp);
    }
#pragma unroll
    for (int r = 0; r < RPT; ++r) rt[r] = yv[r] - tmp[r];
  };
#pragma unroll
  for (int r = 0; r < RPT; ++r) { rr[r] = live ? d.gam[(size_t)(i0 + r) * K + tj] : T(0); xx[r] = T(0); }
  precond();
  T part = T(0);
#pragma unroll
  for (int r = 0; r < RPT; ++r) { pp[r] = rt[r]; part += rr[r] * rt[r]; }
  T nu = bsum(part);
  if (tid == 0 && d.nu_trace) d.nu_trace[(size_t)b * NU_TRACE_LEN] = fabs(nu);
  int iters = 0;
  for (int it = 0; it < max_iter; ++it) {
    const T inv_nu = T(1) / nu;               // off the critical path: beta = nu' / nu needs it only at the end of the iteration
    publish(V, pp);
    __syncthreads();
    T uc[MC], hu = T(0), ap[RPT];
    abt(V, uc);
#pragma unroll
    for (int i = 0; i < MC; ++i) {
      uc[i] = (sval[i] ? V[kV + sidx[i]] : T(0)) - uc[i];
      hu += hh[i] * uc[i];
    }
    hu = rank1 ? sS * quad(hu) : T(0);
#pragma unroll
    for (int i = 0; i < MC; ++i) {
      uc[i] = dinv[i] * uc[i] - hh[i] * hu;
      if (cval[i]) W[kW + c0 + i] = uc[i];
    }
    __syncthreads();
    abmul(uc, W, W, T(-1), ap);
    part = T(0);
#pragma unroll
    for (int r = 0; r < RPT; ++r) part += pp[r] * ap[r];
    const T pAp = bsum(part);
    const T alpha = nu / pAp;
#pragma unroll
    for (int r = 0; r < RPT; ++r) { rr[r] -= ap[r] * alpha; xx[r] += pp[r] * alpha; }
    precond();
    part = T(0);
#pragma unroll
    for (int r = 0; r < RPT; ++r) part += rr[r] * rt[r];
    const T nu_prime = bsum(part);
    iters = it + 1;
    if (tid == 0 && d.nu_trace && iters < NU_TRACE_LEN) d.nu_trace[(size_t)b * NU_TRACE_LEN + iters] = fabs(nu_prime);
    if (fabs(nu_prime) < tol) break;
    const T beta = nu_prime * inv_nu;
#pragma unroll
    for (int r = 0; r < RPT; ++r) pp[r] = rt[r] + pp[r] * beta;
    nu = nu_prime;
  }
  if (live) {
#pragma unroll
    for (int r = 0; r < RPT; ++r) d.l[(size_t)(i0 + r) * K + tj] = xx[r];
  }
  if (tid == 0) {
    d.pcg_iters[b] = iters;
    d.tot_pcg[b] += iters;
    d.tot_qp[b] += 1;
  }
}

// -----------------------------------------------------------------------------------------------------------------
// k_pcg4: k_pcg3's algorithm (matrix-free, register-resident, four lanes per knot, identical arithmetic per product) with the
// shared-memory pipe and the barrier waits taken off the critical path.  ncu on k_pcg3 (profiles/r01_v8_*): the LSU / shared
// data pipe carries ~300 wavefronts per thread and iteration (2 400 cycles per SM and iteration against 1 400 cycles of FP64
// issue and ~4 600 measured), so this variant
//   * splits every product into the part that only needs the group's OWN block (published by the four lanes of the knot, a
//     __syncwarp away) and the part that needs a neighbour's block, and runs the first part BEFORE the block barrier: the barrier
//     wait overlaps AB^T z resp. the AB w products, their shuffles and loads instead of preceding them;
//   * EUL (explicit Euler, tau = 0): AB^T z needs only the velocity half of z (3 instead of 6 128-bit loads);
//   * loads and stores that exist only for the q / qd columns or the top rows are predicated off in the other lanes
//     (W[:, nx:] is never read; full[] only feeds the q rows), which halves their wavefronts.
// Same summation order inside every product as k_pcg3 (the E0 term is contracted differently by the compiler): identical iteration
// counts, solutions equal to rounding (tests/test_gpu_variants.py).  MEASURED: 2.5 % SLOWER than k_pcg3 (18.67 vs 18.19 ns per
// instance-iteration, profiles/README.md) -- the barrier waits it hides are only ~10 % of the stall samples and the extra
// predication costs as much; kept selectable (B2T_PCG_VARIANT=5) as the record of the experiment.
// -----------------------------------------------------------------------------------------------------------------
template <typename T, int MAXT, bool EUL>
__global__ void __launch_bounds__(MAXT, 1) k_pcg4(Dev<T> d, const int* list, const int* count, int stair, T tol, int max_iter) {
  if ((int)blockIdx.x >= *count) return;
  constexpr int LPK = 4;
  constexpr int RPT = (NX % LPK == 0) ? NX / LPK : 1;   // owned rows per lane; divides NJ
  constexpr int MC = (NM + LPK - 1) / LPK;               // Ab columns per lane
  constexpr int NMS = PCG3_NMS;
  using T2 = typename std::conditional<sizeof(T) == 8, double2, float2>::type;
  const int b = list[blockIdx.x];
  const int N = d.N;
  const size_t K = d.K;
  const int tid = threadIdx.x, nt = blockDim.x;
  const bool live = tid < LPK * N;
  const int k = live ? tid / LPK : 0;
  const int g = tid % LPK;
  const bool has_next = live && (k < N - 1);
  const int jo = (k + 1 == N) ? 0 : k + 1;
  const int c0 = g * MC, i0 = g * RPT;
  const size_t tk = (size_t)b * N + k, tj = (size_t)b * N + jo;
  const T dte = d.integrator == 0 ? d.dt : T(0);
  const T tau = d.integrator == 0 ? T(0) : d.dt;
  extern __shared__ unsigned char smem_raw[];
  T* V = reinterpret_cast<T*>(smem_raw);      // [(N+1)][NX]  p / r / O y; block N stays zero
  T* V2 = V + (N + 1) * NX;                   // [(N+1)][NX]  y; block N stays zero
  T* W = V2 + (N + 1) * NX;                   // [N][NMS]     w or q'   (only the first NX entries of a row are ever read)
  T* Wq = W + N * NMS;                        // [N][NMS]     q
  T* red = Wq + N * NMS;                      // 2 x 32 warp partial sums
  T ab[NJ][MC], dinv[MC], hh[MC], pd[RPT][NX];
  bool cval[MC], sval[MC];
  T emul[MC];
  int eidx[MC], sidx[MC];
#pragma unroll
  for (int i = 0; i < MC; ++i) {
    const int c = c0 + i;
    cval[i] = live && c < NM;
    dinv[i] = cval[i] ? d.Gh[(size_t)c * K + tk] : T(0);
    hh[i] = cval[i] ? d.Gh[(size_t)(NM + c) * K + tk] : T(0);
#pragma unroll
    for (int a = 0; a < NJ; ++a)
      ab[a][i] = (cval[i] && has_next) ? d.dt * d.dyn[(size_t)(a * 3 * NJ + c) * K + tk] + ((c == NJ + a) ? T(1) : T(0)) : T(0);
    emul[i] = (!has_next || !cval[i] || c >= NX) ? T(0) : (c < NJ ? T(1) : dte);
    eidx[i] = (c < NJ) ? c : ((c < NX) ? c - NJ : 0);
    sval[i] = cval[i] && c < NX;
    sidx[i] = (c < NX) ? c : 0;
  }
  const T sS = live ? d.Gh[(size_t)(2 * NM) * K + tk] : T(0);
  bool hnz = false;
#pragma unroll
  for (int i = 0; i < MC; ++i) hnz = hnz || (hh[i] != T(0));
  const bool rank1 = d.lim.any != 0 && __any_sync(0xffffffffu, hnz);
#pragma unroll
  for (int r = 0; r < RPT; ++r)
#pragma unroll
    for (int c = 0; c < NX; ++c) pd[r][c] = live ? d.Pd[(size_t)((i0 + r) * NX + c) * K + tj] : T(0);
  for (int idx = tid; idx < 2 * (N + 1) * NX + 2 * N * NMS + 64; idx += nt) V[idx] = T(0);
  __syncthreads();
  const bool top = i0 < NJ;
  const bool odd = (i0 % NJ) != 0;
  const int ownV = jo * NX, ownW = jo * NMS + i0, kV = k * NX, nV = (k + 1) * NX, kW = k * NMS;

  auto quad = [&](T v) -> T {
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    return v;
  };
  int red_sel = 0;
  auto bsum = [&](T v) -> T {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    T* rb = red + 32 * red_sel;
    red_sel ^= 1;
    if ((tid & 31) == 0) rb[tid >> 5] = v;
    __syncthreads();
    constexpr int NW = MAXT / 32;
    T t[NW];
#pragma unroll
    for (int i = 0; i < NW; i += 2) { const T2 p2 = *reinterpret_cast<const T2*>(rb + i); t[i] = p2.x; t[i + 1] = p2.y; }
#pragma unroll
    for (int st = 1; st < NW; st *= 2)
#pragma unroll
      for (int i = 0; i + st < NW; i += 2 * st) t[i] += t[i + st];
    return t[0];
  };
  auto publish = [&](T* buf, const T* val) {
    if (live) {
#pragma unroll
      for (int r = 0; r < RPT; ++r) buf[ownV + i0 + r] = val[r];
    }
  };
  // tc_i = (AB_k^T z)_c for my columns, z = block k+1 of `buf` = the group's own block (or the zero block for k = N-1)
  auto abt = [&](const T* buf, T* tc) {
    T pi[NJ];
#pragma unroll
    for (int a = 0; a < NJ; a += 2) {
      const T2 zv = *reinterpret_cast<const T2*>(buf + nV + NJ + a);
      if constexpr (EUL) {
        pi[a] = zv.x;
        if (a + 1 < NJ) pi[a + 1] = zv.y;
      } else {
        const T2 zq = *reinterpret_cast<const T2*>(buf + nV + a);
        pi[a] = tau * zq.x + zv.x;
        if (a + 1 < NJ) pi[a + 1] = tau * zq.y + zv.y;
      }
    }
#pragma unroll
    for (int i = 0; i < MC; ++i) {
      T acc0 = (emul[i] != T(0)) ? emul[i] * buf[nV + eidx[i]] : T(0), acc1 = T(0);
#pragma unroll
      for (int a = 0; a < NJ; a += 2) {
        acc0 += ab[a][i] * pi[a];
        if (a + 1 < NJ) acc1 += ab[a + 1][i] * pi[a + 1];
      }
      tc[i] = acc0 + acc1;
    }
  };
  // first half of  out = AB_k zz + sign Wn[jo]:  val_r = (AB_k zz)_row for the owned rows; zz: my columns in zc, all columns in `full`
  // (written by the lanes of this group: a __syncwarp away)
  auto abmul_own = [&](const T* zc, const T* full, T* val) {
    T pb[NJ];
#pragma unroll
    for (int a = 0; a < NJ; ++a) {
      T acc = T(0);
#pragma unroll
      for (int i = 0; i < MC; ++i) acc += ab[a][i] * zc[i];
      pb[a] = acc;
    }
    T bot[RPT];
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
      const T send = odd ? pb[r] : pb[RPT + r];
      T keep = odd ? pb[RPT + r] : pb[r];
      keep += __shfl_xor_sync(0xffffffffu, send, 1);
      keep += __shfl_xor_sync(0xffffffffu, keep, 2);
      bot[r] = keep;
    }
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
      if (top) {
        const T zq = full[kW + i0 + r], zd = full[kW + NJ + i0 + r];
        val[r] = zq + dte * zd + tau * bot[r];
      } else {
        val[r] = bot[r];
      }
    }
  };
  // second half, after the block barrier: the neighbour's rows
  auto abmul_nb = [&](const T* val, const T* Wn, T sign, T* out) {
#pragma unroll
    for (int r = 0; r < RPT; ++r) out[r] = has_next ? val[r] + sign * Wn[ownW + r] : (live ? sign * Wn[ownW + r] : T(0));
  };
  auto pd_mul = [&](const T* buf, T* out) {
    T o0[RPT], o1[RPT];
#pragma unroll
    for (int r = 0; r < RPT; ++r) { o0[r] = T(0); o1[r] = T(0); }
#pragma unroll
    for (int c = 0; c < NX; c += 2) {
      const T2 v = *reinterpret_cast<const T2*>(buf + ownV + c);
#pragma unroll
      for (int r = 0; r < RPT; ++r) { o0[r] += pd[r][c] * v.x; o1[r] += pd[r][c + 1] * v.y; }
    }
#pragma unroll
    for (int r = 0; r < RPT; ++r) out[r] = o0[r] + o1[r];
  };
  T rr[RPT], xx[RPT], pp[RPT], rt[RPT], yv[RPT], tmp[RPT], val[RPT];
  auto precond = [&]() {
    publish(V, rr);
    __syncwarp();
    pd_mul(V, yv);
    if (!stair) {
#pragma unroll
      for (int r = 0; r < RPT; ++r) rt[r] = yv[r];
      return;
    }
    publish(V2, yv);
    __syncwarp();
    T u1[MC], u2[MC], h1 = T(0), h2 = T(0);
    abt(V2, u2);                                // own block: before the barrier
    __syncthreads();
#pragma unroll
    for (int i = 0; i < MC; ++i) {
      u1[i] = sval[i] ? V2[kV + sidx[i]] : T(0);
      h1 += hh[i] * u1[i];
      h2 += hh[i] * u2[i];
    }
    if (rank1) { h1 = sS * quad(h1); h2 = sS * quad(h2); } else { h1 = T(0); h2 = T(0); }
#pragma unroll
    for (int i = 0; i < MC; ++i) {
      u1[i] = dinv[i] * u1[i] - hh[i] * h1;
      u2[i] = dinv[i] * u2[i] - hh[i] * h2;
      if (sval[i]) { Wq[kW + c0 + i] = u1[i]; W[kW + c0 + i] = u2[i]; }
    }
    __syncwarp();
    abmul_own(u1, Wq, val);                     // own knot: before the barrier
    __syncthreads();
    abmul_nb(val, W, T(1), tmp);
    publish(V, tmp);
    __syncwarp();
    pd_mul(V, tmp);
#pragma unroll
    for (int r = 0; r < RPT; ++r) rt[r] = yv[r] - tmp[r];
  };
#pragma unroll
  for (int r = 0; r < RPT; ++r) { rr[r] = live ? d.gam[(size_t)(i0 + r) * K + tj] : T(0); xx[r] = T(0); }
  precond();
  T part = T(0);
#pragma unroll
  for (int r = 0; r < RPT; ++r) { pp[r] = rt[r]; part += rr[r] * rt[r]; }
  T nu = bsum(part);
  if (tid == 0 && d.nu_trace) d.nu_trace[(size_t)b * NU_TRACE_LEN] = fabs(nu);
  int iters = 0;
  for (int it = 0; it < max_iter; ++it) {
    const T inv_nu = T(1) / nu;
    publish(V, pp);
    __syncwarp();
    T uc[MC], hu = T(0), ap[RPT];
    abt(V, uc);                                 // own block: before the barrier
    __syncthreads();
#pragma unroll
    for (int i = 0; i < MC; ++i) {
      uc[i] = (sval[i] ? V[kV + sidx[i]] : T(0)) - uc[i];
      hu += hh[i] * uc[i];
    }
    hu = rank1 ? sS * quad(hu) : T(0);
#pragma unroll
    for (int i = 0; i < MC; ++i) {
      uc[i] = dinv[i] * uc[i] - hh[i] * hu;
      if (sval[i]) W[kW + c0 + i] = uc[i];
    }
    __syncwarp();
    abmul_own(uc, W, val);
    __syncthreads();
    abmul_nb(val, W, T(-1), ap);
    part = T(0);
#pragma unroll
    for (int r = 0; r < RPT; ++r) part += pp[r] * ap[r];
    const T pAp = bsum(part);
    const T alpha = nu / pAp;
#pragma unroll
    for (int r = 0; r < RPT; ++r) { rr[r] -= ap[r] * alpha; xx[r] += pp[r] * alpha; }
    precond();
    part = T(0);
#pragma unroll
    for (int r = 0; r < RPT; ++r) part += rr[r] * rt[r];
    const T nu_prime = bsum(part);
    iters = it + 1;
    if (tid == 0 && d.nu_trace && iters < NU_TRACE_LEN) d.nu_trace[(size_t)b * NU_TRACE_LEN + iters] = fabs(nu_prime);
    if (fabs(nu_prime) < tol) break;
    const T beta = nu_prime * inv_nu;
#pragma unroll
    for (int r = 0; r < RPT; ++r) pp[r] = rt[r] + pp[r] * beta;
    nu = nu_prime;
  }
  if (live) {
#pragma unroll
    for (int r = 0; r < RPT; ++r) d.l[(size_t)(i0 + r) * K + tj] = xx[r];
  }
  if (tid == 0) {
    d.pcg_iters[b] = iters;
    d.tot_pcg[b] += iters;
    d.tot_qp[b] += 1;
  }
}

// -----------------------------------------------------------------------------------------------------------------
// k_pcg6: the matrix-free PCG of k_pcg3 with SIX lanes per knot (robots with nj % 6 == 0: arm6).  Lane g of a knot holds the three
// columns [3g, 3g+3) of Ab_k -- with nm = 18 columns and six lanes every lane's columns lie in ONE of the blocks q / qd / u, so the
// E0 coupling, the self term and the "only the state part of w is ever read" predicates are lane-uniform scalars instead of
// per-column tables -- and rows [2g, 2g+2) of the preconditioner block of the block row it owns.  Five knots per warp (30 lanes),
// 13 warps for N = 64: the per-thread instruction stream of one PCG iteration is ~35 % shorter than with four lanes and the SM
// runs 13 instead of 8 warps against the same shared-memory / FP64 work, which is what a latency-bound kernel needs (ncu on
// k_pcg3: 6 cycles between issues of a warp, 2 warps per scheduler).  48 doubles of matrix data per lane => <= 152 registers.
// Summation orders differ from k_pcg3 (6-lane reductions) -- same recurrence and exit test (PCG.py:66-111).
// MEASURED: 1.8x SLOWER than k_pcg3 (32.6 vs 18.2 ns per instance-iteration; profiles/r02_v6lane_*).  13 warps put 4 warps on one
// scheduler, whose 16 K registers then allow only 128 per thread: 27 spilled doubles (LDL/STL 5.6 % of the instructions), address and
// constant recomputation (IMAD 12.5 %), divergent predication (BSSY/BSYNC/BRA 10 %); the FP64 instruction count per SM is unchanged,
// so the shorter per-thread stream the design was after never materialises.  Kept selectable (B2T_PCG_VARIANT=6) as the record.
// -----------------------------------------------------------------------------------------------------------------
constexpr int PCG6_KPW = 5;                       // knots per warp
constexpr int PCG6_THREADS = 416;                 // 13 warps: N <= 65
constexpr int PCG6_VS = NX + 2;                   // row stride of the shared vectors: 5 knots x 16-byte accesses fall into distinct banks
template <typename T, bool EUL>
__global__ void __launch_bounds__(PCG6_THREADS, 1) k_pcg6(Dev<T> d, const int* list, const int* count, int stair, T tol, int max_iter) {
  if ((int)blockIdx.x >= *count) return;
  constexpr int LPK = 6;
  constexpr int RPT = (NX % LPK == 0) ? NX / LPK : 1;
  constexpr int MC = (NM % LPK == 0) ? NM / LPK : 1;
  constexpr int VS = PCG6_VS;
  constexpr int NW = PCG6_THREADS / 32;
  using T2 = typename std::conditional<sizeof(T) == 8, double2, float2>::type;
  const int b = list[blockIdx.x];
  const int N = d.N;
  const size_t K = d.K;
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5;
  const int kq = lane / LPK, g = lane - kq * LPK;
  const bool valid = lane < PCG6_KPW * LPK;
  const int kk = warp * PCG6_KPW + kq;
  const bool live = valid && kk < N;
  const int k = live ? kk : 0;
  const bool has_next = live && (k < N - 1);
  const int jo = (k + 1 == N) ? 0 : k + 1;
  const int cls = g / 2;                      // 0: q columns, 1: qd columns, 2: u columns (MC = NJ / 2)
  const bool xcol = live && cls < 2;          // my columns belong to the state part
  const bool top = g < LPK / 2;               // my rows are q rows of [A B]
  const int j3 = g % 3;                       // position inside the half
  const int c0 = g * MC, i0 = g * RPT;
  const int base = valid ? kq * LPK : lane;   // first lane of my knot group (idle lanes talk to themselves)
  const int partner = valid ? (top ? lane + 3 : lane - 3) : lane;
  const int src1 = valid ? base + (top ? 0 : 3) + (j3 + 2) % 3 : lane;     // the lane one position before me in my half
  const int src2 = valid ? base + (top ? 0 : 3) + (j3 + 1) % 3 : lane;     // two positions before me
  const size_t tk = (size_t)b * N + k, tj = (size_t)b * N + jo;
  const T dte = d.integrator == 0 ? d.dt : T(0);
  const T tau = d.integrator == 0 ? T(0) : d.dt;
  extern __shared__ unsigned char smem_raw[];
  T* V = reinterpret_cast<T*>(smem_raw);      // [(N+1)][VS]  p / r / O y; block N stays zero
  T* V2 = V + (N + 1) * VS;                   // [(N+1)][VS]  y
  T* W = V2 + (N + 1) * VS;                   // [N][VS]      state part of w or q'
  T* Wq = W + N * VS;                         // [N][VS]      state part of q
  T* red = Wq + N * VS;                       // 2 x 16 warp partial sums
  T ab[NJ][MC], dinv[MC], hh[MC], pd[RPT][NX];
#pragma unroll
  for (int i = 0; i < MC; ++i) {
    const int c = c0 + i;
    dinv[i] = live ? d.Gh[(size_t)c * K + tk] : T(0);
    hh[i] = live ? d.Gh[(size_t)(NM + c) * K + tk] : T(0);
#pragma unroll
    for (int a = 0; a < NJ; ++a)
      ab[a][i] = has_next ? d.dt * d.dyn[(size_t)(a * 3 * NJ + c) * K + tk] + ((c == NJ + a) ? T(1) : T(0)) : T(0);
  }
  const T ecoef = has_next ? (cls == 0 ? T(1) : (cls == 1 ? dte : T(0))) : T(0);     // E0^T z_q: column c takes z_q[c % nj]
  const int ec0 = (cls == 1) ? c0 - NJ : (cls == 0 ? c0 : 0);
  const T sS = live ? d.Gh[(size_t)(2 * NM) * K + tk] : T(0);
  bool hnz = false;
#pragma unroll
  for (int i = 0; i < MC; ++i) hnz = hnz || (hh[i] != T(0));
  const bool rank1 = d.lim.any != 0 && __any_sync(0xffffffffu, hnz);
#pragma unroll
  for (int r = 0; r < RPT; ++r)
#pragma unroll
    for (int c = 0; c < NX; ++c) pd[r][c] = live ? d.Pd[(size_t)((i0 + r) * NX + c) * K + tj] : T(0);
  for (int idx = tid; idx < 2 * (N + 1) * VS + 2 * N * VS + 32; idx += nt) V[idx] = T(0);
  __syncthreads();
  const int ownV = jo * VS, ownW = jo * VS + i0, kV = k * VS, nV = (k + 1) * VS, kW = k * VS;

  // sum over the six lanes of the knot group, identical bits in all of them: pairs, then the three pair sums in a fixed order
  auto gsum = [&](T v) -> T {
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    const T p0 = __shfl_sync(0xffffffffu, v, base), p1 = __shfl_sync(0xffffffffu, v, valid ? base + 2 : lane),
            p2 = __shfl_sync(0xffffffffu, v, valid ? base + 4 : lane);
    return (p0 + p1) + p2;
  };
  // deterministic block sum: warp butterfly, one barrier, butterfly over the (<= 16) warp partials -- the same pairwise tree in every lane
  int red_sel = 0;
  auto bsum = [&](T v) -> T {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    T* rb = red + 16 * red_sel;
    red_sel ^= 1;
    if (lane == 0) rb[warp] = v;
    __syncthreads();
    T t = rb[lane & 15];
#pragma unroll
    for (int o = 1; o < 16; o <<= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    return t;
  };
  static_assert(NW <= 16, "bsum holds 16 warp partials");
  auto publish = [&](T* buf, const T* val) {
    if (live) {
#pragma unroll
      for (int r = 0; r < RPT; ++r) buf[ownV + i0 + r] = val[r];
    }
  };
  // tc_i = (AB_k^T z)_c for my columns; z = block k+1 of `buf` = the group's own block (zero block for k = N-1)
  auto abt = [&](const T* buf, T* tc) {
    T pi[NJ];
#pragma unroll
    for (int a = 0; a < NJ; a += 2) {
      const T2 zv = *reinterpret_cast<const T2*>(buf + nV + NJ + a);
      if constexpr (EUL) {
        pi[a] = zv.x;
        if (a + 1 < NJ) pi[a + 1] = zv.y;
      } else {
        const T2 zq = *reinterpret_cast<const T2*>(buf + nV + a);
        pi[a] = tau * zq.x + zv.x;
        if (a + 1 < NJ) pi[a + 1] = tau * zq.y + zv.y;
      }
    }
#pragma unroll
    for (int i = 0; i < MC; ++i) {
      T acc0 = xcol ? ecoef * buf[nV + ec0 + i] : T(0), acc1 = T(0);
#pragma unroll
      for (int a = 0; a < NJ; a += 2) {
        acc0 += ab[a][i] * pi[a];
        if (a + 1 < NJ) acc1 += ab[a + 1][i] * pi[a + 1];
      }
      tc[i] = acc0 + acc1;
    }
  };
  // val_r = (AB_k zz)_row for my rows; zz: my columns in zc (registers), its state part in `full` (written by this group)
  auto abmul_own = [&](const T* zc, const T* full, T* val) {
    T pb[NJ];
#pragma unroll
    for (int a = 0; a < NJ; ++a) {
      T acc = T(0);
#pragma unroll
      for (int i = 0; i < MC; ++i) acc += ab[a][i] * zc[i];
      pb[a] = acc;
    }
    // reduce-scatter over the six lanes: rows (2 j3, 2 j3 + 1) of the half sums go to position j3 of each half (two rounds inside the
    // half, the sender picks what its receiver needs), then the two halves are added
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
      // what the lane one / two positions AFTER me needs: its rows RPT * ((j3 + 1) % 3) + r, RPT * ((j3 + 2) % 3) + r
      const T s1 = j3 == 0 ? pb[RPT + r] : (j3 == 1 ? pb[2 * RPT + r] : pb[r]);
      const T s2 = j3 == 0 ? pb[2 * RPT + r] : (j3 == 1 ? pb[r] : pb[RPT + r]);
      T keep = j3 == 0 ? pb[r] : (j3 == 1 ? pb[RPT + r] : pb[2 * RPT + r]);
      keep += __shfl_sync(0xffffffffu, s1, src1);
      keep += __shfl_sync(0xffffffffu, s2, src2);
      keep += __shfl_sync(0xffffffffu, keep, partner);
      if (top) {
        const T zq = full[kW + i0 + r], zd = full[kW + NJ + i0 + r];
        val[r] = zq + dte * zd + tau * keep;
      } else {
        val[r] = keep;
      }
    }
  };
  auto abmul_nb = [&](const T* val, const T* Wn, T sign, T* out) {
#pragma unroll
    for (int r = 0; r < RPT; ++r) out[r] = has_next ? val[r] + sign * Wn[ownW + r] : (live ? sign * Wn[ownW + r] : T(0));
  };
  auto pd_mul = [&](const T* buf, T* out) {
    T o0[RPT], o1[RPT];
#pragma unroll
    for (int r = 0; r < RPT; ++r) { o0[r] = T(0); o1[r] = T(0); }
#pragma unroll
    for (int c = 0; c < NX; c += 2) {
      const T2 v = *reinterpret_cast<const T2*>(buf + ownV + c);
#pragma unroll
      for (int r = 0; r < RPT; ++r) { o0[r] += pd[r][c] * v.x; o1[r] += pd[r][c + 1] * v.y; }
    }
#pragma unroll
    for (int r = 0; r < RPT; ++r) out[r] = o0[r] + o1[r];
  };
  T rr[RPT], xx[RPT], pp[RPT], rt[RPT], yv[RPT], tmp[RPT], val[RPT];
  auto precond = [&]() {
    publish(V, rr);
    __syncwarp();
    pd_mul(V, yv);
    if (!stair) {
#pragma unroll
      for (int r = 0; r < RPT; ++r) rt[r] = yv[r];
      return;
    }
    publish(V2, yv);
    __syncwarp();
    T u1[MC], u2[MC], h1 = T(0), h2 = T(0);
    abt(V2, u2);                                // own block: before the barrier
    __syncthreads();
#pragma unroll
    for (int i = 0; i < MC; ++i) {
      u1[i] = xcol ? V2[kV + c0 + i] : T(0);
      h1 += hh[i] * u1[i];
      h2 += hh[i] * u2[i];
    }
    if (rank1) { h1 = sS * gsum(h1); h2 = sS * gsum(h2); } else { h1 = T(0); h2 = T(0); }
#pragma unroll
    for (int i = 0; i < MC; ++i) {
      u1[i] = dinv[i] * u1[i] - hh[i] * h1;
      u2[i] = dinv[i] * u2[i] - hh[i] * h2;
      if (xcol) { Wq[kW + c0 + i] = u1[i]; W[kW + c0 + i] = u2[i]; }
    }
    __syncwarp();
    abmul_own(u1, Wq, val);                     // own knot: before the barrier
    __syncthreads();
    abmul_nb(val, W, T(1), tmp);
    publish(V, tmp);
    __syncwarp();
    pd_mul(V, tmp);
#pragma unroll
    for (int r = 0; r < RPT; ++r) rt[r] = yv[r] - tmp[r];
  };
#pragma unroll
  for (int r = 0; r < RPT; ++r) { rr[r] = live ? d.gam[(size_t)(i0 + r) * K + tj] : T(0); xx[r] = T(0); }
  precond();
  T part = T(0);
#pragma unroll
  for (int r = 0; r < RPT; ++r) { pp[r] = rt[r]; part += rr[r] * rt[r]; }
  T nu = bsum(part);
  if (tid == 0 && d.nu_trace) d.nu_trace[(size_t)b * NU_TRACE_LEN] = fabs(nu);
  int iters = 0;
  for (int it = 0; it < max_iter; ++it) {
    const T inv_nu = T(1) / nu;
    publish(V, pp);
    __syncwarp();
    T uc[MC], hu = T(0), ap[RPT];
    abt(V, uc);                                 // own block: before the barrier
    __syncthreads();
#pragma unroll
    for (int i = 0; i < MC; ++i) {
      uc[i] = (xcol ? V[kV + c0 + i] : T(0)) - uc[i];
      hu += hh[i] * uc[i];
    }
    hu = rank1 ? sS * gsum(hu) : T(0);
#pragma unroll
    for (int i = 0; i < MC; ++i) {
      uc[i] = dinv[i] * uc[i] - hh[i] * hu;
      if (xcol) W[kW + c0 + i] = uc[i];
    }
    __syncwarp();
    abmul_own(uc, W, val);
    __syncthreads();
    abmul_nb(val, W, T(-1), ap);
    part = T(0);
#pragma unroll
    for (int r = 0; r < RPT; ++r) part += pp[r] * ap[r];
    const T pAp = bsum(part);
    const T alpha = nu / pAp;
#pragma unroll
    for (int r = 0; r < RPT; ++r) { rr[r] -= ap[r] * alpha; xx[r] += pp[r] * alpha; }
    precond();
    part = T(0);
#pragma unroll
    for (int r = 0; r < RPT; ++r) part += rr[r] * rt[r];
    const T nu_prime = bsum(part);
    iters = it + 1;
    if (tid == 0 && d.nu_trace && iters < NU_TRACE_LEN) d.nu_trace[(size_t)b * NU_TRACE_LEN + iters] = fabs(nu_prime);
    if (fabs(nu_prime) < tol) break;
    const T beta = nu_prime * inv_nu;
#pragma unroll
    for (int r = 0; r < RPT; ++r) pp[r] = rt[r] + pp[r] * beta;
    nu = nu_prime;
  }
  if (live) {
#pragma unroll
    for (int r = 0; r < RPT; ++r) d.l[(size_t)(i0 + r) * K + tj] = xx[r];
  }
  if (tid == 0) {
    d.pcg_iters[b] = iters;
    d.tot_pcg[b] += iters;
    d.tot_qp[b] += 1;
  }
}

// -----------------------------------------------------------------------------------------------------------------
// k_bt_solve: exact solve of S l = gamma (methods 'S' and 'N' of the reference: np.linalg.solve on the Schur system,
// TrajoptMPCReference.py:430-436, resp. on the full KKT system :349-357 -- the same solution up to rounding).
// Block Thomas algorithm on the SPD block-tridiagonal matrix T = -S, one thread per instance:
//   Delta_0 = D_0,  Delta_j = D_j - O_j Delta_{j-1}^-1 O_j^T,  g_j = b_j - O_j Delta_{j-1}^-1 g_{j-1},
//   l_{N-1} = Delta_{N-1}^-1 g_{N-1},  l_j = Delta_j^-1 (g_j - O_{j+1}^T l_{j+1}).
// Delta_j^-1 is kept in the (otherwise unused) preconditioner array, g_j in l.  ~5 k FMA per knot: far less work than PCG,
// so a latency-bound kernel is adequate here.
// -----------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(32) k_bt_solve(Dev<T> d, const int* list, const int* count) {
  const int slot = blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= *count) return;
  const int b = list[slot];
  const int N = d.N;
  const size_t K = d.K;
  const size_t t0 = (size_t)b * N;
  T Dp[NX * NX], Dj[NX * NX], M[NX * NX], gp[NX], gj[NX];
  for (int j = 0; j < N; ++j) {
    const size_t t = t0 + j;
    for (int e = 0; e < NX * NX; ++e) Dj[e] = -d.Sd[(size_t)e * K + t];
    for (int i = 0; i < NX; ++i) gj[i] = -d.gam[(size_t)i * K + t];
    if (j > 0) {
      // M = O_j Dp   (O_j = -So_j)
      for (int i = 0; i < NX; ++i)
        for (int c = 0; c < NX; ++c) {
          T acc = T(0);
          for (int r = 0; r < NX; ++r) acc += (-d.So[(size_t)(i * NX + r) * K + t]) * Dp[r * NX + c];
          M[i * NX + c] = acc;
        }
      for (int i = 0; i < NX; ++i) {
        for (int c = 0; c < NX; ++c) {
          T acc = T(0);
          for (int r = 0; r < NX; ++r) acc += M[i * NX + r] * (-d.So[(size_t)(c * NX + r) * K + t]);
          Dj[i * NX + c] -= acc;
        }
        T accg = T(0);
        for (int r = 0; r < NX; ++r) accg += M[i * NX + r] * gp[r];
        gj[i] -= accg;
      }
    }
    spd_inverse_inplace(Dj, NX, NX);
    for (int e = 0; e < NX * NX; ++e) { Dp[e] = Dj[e]; d.Pd[(size_t)e * K + t] = Dj[e]; }
    for (int i = 0; i < NX; ++i) { gp[i] = gj[i]; d.l[(size_t)i * K + t] = gj[i]; }
  }
  // back substitution
  T ln[NX];
  for (int j = N - 1; j >= 0; --j) {
    const size_t t = t0 + j;
    T rhs[NX];
    for (int i = 0; i < NX; ++i) rhs[i] = d.l[(size_t)i * K + t];
    if (j < N - 1) {
      for (int i = 0; i < NX; ++i) {
        T acc = T(0);
        for (int r = 0; r < NX; ++r) acc += (-d.So[(size_t)(r * NX + i) * K + t + 1]) * ln[r];     // O_{j+1}^T l_{j+1}
        rhs[i] -= acc;
      }
    }
    T lj[NX];
    for (int i = 0; i < NX; ++i) {
      T acc = T(0);
      for (int c = 0; c < NX; ++c) acc += d.Pd[(size_t)(i * NX + c) * K + t] * rhs[c];
      lj[i] = acc;
    }
    for (int i = 0; i < NX; ++i) { ln[i] = lj[i]; d.l[(size_t)i * K + t] = lj[i]; }
  }
  d.pcg_iters[b] = 0;
  d.tot_qp[b] += 1;
}

// -----------------------------------------------------------------------------------------------------------------
// k_recover: dz_k = Ghat_k (g_k - [l_k; 0] + AB_k^T l_{k+1})     (:449-452)
// -----------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(128) k_recover(Dev<T> d, const int* list, const int* count) {
  const size_t gt = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int slot = (int)(gt / d.N);
  if (slot >= *count) return;
  const int k = (int)(gt % d.N);
  const int b = list[slot];
  const size_t t = (size_t)b * d.N + k;
  const size_t K = d.K;
  (void)0;
  const bool terminal = (k == d.N - 1);
  T rhs[NM];
  for (int i = 0; i < NM; ++i) rhs[i] = d.g[(size_t)i * K + t];
  const T* l = d.l + (size_t)b * d.N;
  for (int i = 0; i < NX; ++i) rhs[i] -= l[(size_t)i * K + k];
  if (!terminal) {
    T AB[NX * NM];
    load_AB(d, t, AB);
    for (int c = 0; c < NM; ++c) {
      T acc = T(0);
      for (int i = 0; i < NX; ++i) acc += AB[i * NM + c] * l[(size_t)i * K + k + 1];
      rhs[c] += acc;
    }
  }
  T z[NM];
  if (d.hard.any) load_xu(d.x, d.u, K, t, terminal, z, z + NX);
  for (int i = 0; i < NM; ++i) {
    T acc = T(0);
    for (int c = 0; c < NM; ++c) acc += d.Gh[(size_t)(i * NM + c) * K + t] * rhs[c];
    T tf = T(0);
    if (d.hard.any && i < (terminal ? NX : NM) && hard_active(d.hard, z, i, &tf)) acc += tf;      // dz = Ghat' (g' - C^T l) + E_A t  (k_kkt)
    d.dz[(size_t)i * K + t] = acc;
  }
}

// -----------------------------------------------------------------------------------------------------------------
// merit evaluation of one instance (block): J = totalCost (:296-310), c = totalHardConstraintViolation (:273-294),
// D = directional derivative (:635-648).  Per-knot terms in parallel, then summed by thread 0 in the reference's
// sequential order.  TRIAL selects (xn, un, xkp1n) instead of (x, u, xkp1).
// smem: 6 * N scalars.
// -----------------------------------------------------------------------------------------------------------------
template <typename T, bool TRIAL, bool WITH_D, bool WITH_C>
__device__ __forceinline__ void merit_terms(const Dev<T>& d, int b, T* sm, T* J_out, T* c_out, T* D_out) {
  const int N = d.N;
  const size_t K = d.K;
  T* s_cost = sm; T* s_soft = sm + N; T* s_c = sm + 2 * N; T* s_D = sm + 3 * N; T* s_Ds = sm + 4 * N; T* s_h = sm + 5 * N;
  const T* X = TRIAL ? d.xn : d.x;
  const T* U = TRIAL ? d.un : d.u;
  const T* XK = TRIAL ? d.xkp1n : d.xkp1;
  for (int k = threadIdx.x; k < N; k += blockDim.x) {
    const size_t t = (size_t)b * N + k;
    const bool terminal = (k == N - 1);
    T z[NM], xg[NX];
    load_xu(X, U, K, t, terminal, z, z + NX);
    for (int i = 0; i < NX; ++i) xg[i] = d.xg[(size_t)i * d.B + b];
    s_cost[k] = cost_value(d.cost, z, z + NX, xg, k, terminal);
    s_soft[k] = d.lim.any ? soft_value(d.lim, z, d.mu + t, d.lam + t, K, terminal) : T(0);
    if constexpr (WITH_C) {
      T acc = T(0);
      if (k == 0) {
        for (int i = 0; i < NX; ++i) acc += fabs(z[i] - d.xs[(size_t)i * d.B + b]);
      } else {
        for (int i = 0; i < NX; ++i) acc += fabs(z[i] - XK[(size_t)i * K + t - 1]);
      }
      s_c[k] = acc;
      if (d.hard.any) s_h[k] = hard_violation(d.hard, z, terminal);
    }
    if constexpr (WITH_D) {
      T g[NM], dzk[NM];
      cost_grad_hess<T, false>(d.cost, z, z + NX, xg, k, terminal, g, (T*)nullptr);
      const int M = terminal ? NX : NM;
      for (int i = 0; i < NM; ++i) dzk[i] = d.dz[(size_t)i * K + t];
      T acc = T(0);
      for (int i = 0; i < M; ++i) acc += g[i] * dzk[i];
      s_D[k] = acc;
      T accs = T(0);
      if (d.lim.any) {
        T gck[NM];
        soft_grad(d.lim, z, d.mu + t, d.lam + t, K, terminal, gck);
        for (int i = 0; i < M; ++i) accs += gck[i] * dzk[i];
      }
      s_Ds[k] = accs;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    T J = T(0);
    for (int k = 0; k < N; ++k) J += s_cost[k];
    if (d.lim.any)
      for (int k = 0; k < N; ++k) J += s_soft[k];
    *J_out = J;
    if constexpr (WITH_C) {
      T c = T(0);
      for (int k = 0; k < N; ++k) c += s_c[k];
      if (d.hard.any)
        for (int k = 0; k < N; ++k) c += s_h[k];      // violated hard bounds, after the dynamics defects (:285-293)
      *c_out = c;
    }
    if constexpr (WITH_D) {
      T D = T(0);
      for (int k = 0; k < N; ++k) {
        D += s_D[k];
        if (d.lim.any) D += s_Ds[k];
      }
      *D_out = D;
    }
  }
  __syncthreads();
}

template <typename T>
__device__ __forceinline__ void trace_row(const Dev<T>& d, int b, int ls, T alpha, T D, T ratio, int inner, int success) {
  int row = d.trace_rows[b];
  if (row < d.trace_cap) {
    T* tr = d.trace + ((size_t)b * d.trace_cap + row) * TRACE_FIELDS;
    tr[0] = (T)d.outer_iter[b]; tr[1] = (T)d.sqp_iter[b]; tr[2] = (T)ls; tr[3] = alpha; tr[4] = d.rho[b]; tr[5] = d.J[b];
    tr[6] = d.c[b]; tr[7] = d.merit[b]; tr[8] = D; tr[9] = ratio; tr[10] = (T)inner; tr[11] = (T)success;
  }
  d.trace_rows[b] = row + 1;
}

// start of an outer (soft-constraint) iteration: rho, J, c, merit and the seed trace row (SQP :535-569)
template <typename T>
__global__ void k_outer_begin(Dev<T> d, const int* list, const int* count, Opts<T> o, int with_c) {
  if ((int)blockIdx.x >= *count) return;
  const int b = list[blockIdx.x];
  extern __shared__ unsigned char smem_raw[];
  T* sm = reinterpret_cast<T*>(smem_raw);
  __shared__ T J, c, D;
  if (with_c) merit_terms<T, false, false, true>(d, b, sm, &J, &c, &D);
  else merit_terms<T, false, false, false>(d, b, sm, &J, &c, &D);
  if (threadIdx.x == 0) {
    if (with_c) d.c[b] = c;
    d.J[b] = J;
    d.rho[b] = o.rho_init;
    d.drho[b] = T(1);
    d.merit[b] = J + o.merit_mu * d.c[b];
    d.sqp_iter[b] = 0;
    d.trace_rows[b] = 0;
    d.phase[b] = PH_SQP;
    trace_row(d, b, 0, T(1), T(0), T(0), 0, 0);
  }
}

// start of one SQP iteration for all active instances: line-search state
template <typename T>
__global__ void k_iter_begin(Dev<T> d) {
  const int n = *d.n_act;
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s == 0) {
    d.n_ls[0] = n;
    for (int i = 1; i <= MAX_LS_TRIALS; ++i) d.n_ls[i] = 0;
    *d.n_restart = 0;
  }
  if (s >= n) return;
  const int b = d.act[s];
  d.alpha[b] = T(1);
  d.ls_iter[b] = 0;
  d.err[b] = 0;
  d.ls_list0[s] = b;
}

// -----------------------------------------------------------------------------------------------------------------
// k_merit: evaluates the trial point of every instance still in its line search and takes the accept / backtrack /
// fail decision (SQP :629-744, reduce_regularization :457-461).
// -----------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void k_merit(Dev<T> d, const int* list, const int* count, int* next_list, int* next_count, Opts<T> o) {
  if ((int)blockIdx.x >= *count) return;
  const int b = list[blockIdx.x];
  extern __shared__ unsigned char smem_raw[];
  T* sm = reinterpret_cast<T*>(smem_raw);
  __shared__ T Jn, cn, D;
  __shared__ int accept;
  merit_terms<T, true, true, true>(d, b, sm, &Jn, &cn, &D);
  if (threadIdx.x == 0) {
    const T mu = o.merit_mu;
    const T alpha = d.alpha[b];
    const T merit_new = Jn + mu * cn;
    const T delta_J = d.J[b] - Jn;
    const T delta_merit = d.merit[b] - merit_new;
    const T expected = alpha * (D - mu * cn);
    const T ratio = delta_merit / expected;
    d.deltaJ[b] = delta_J;
    d.tot_trials[b] += 1;
    accept = 0;
    if (delta_merit >= T(0) && ratio >= o.er_min && ratio <= o.er_max) {
      accept = 1;
      d.J[b] = Jn; d.c[b] = cn; d.merit[b] = merit_new;
      T drho = fmin(d.drho[b] / o.rho_factor, T(1) / o.rho_factor);
      T rho = fmax(d.rho[b] * drho, o.rho_min);
      d.drho[b] = drho; d.rho[b] = rho;
      trace_row(d, b, d.ls_iter[b], alpha, D, ratio, d.pcg_iters[b], 1);
    } else if (alpha > o.alpha_min) {
      d.alpha[b] = alpha * o.alpha_factor;
      d.ls_iter[b] += 1;
      const int pos = atomicAdd(next_count, 1);
      next_list[pos] = b;
    } else {
      d.err[b] = 1;
      trace_row(d, b, d.ls_iter[b], alpha, D, ratio, d.pcg_iters[b], 0);
    }
  }
  __syncthreads();
  if (accept) {
    const int N = d.N;
    const size_t K = d.K;
    for (int idx = threadIdx.x; idx < N * NX; idx += blockDim.x) {
      const int i = idx / N, k = idx % N;
      const size_t t = (size_t)b * N + k;
      d.x[(size_t)i * K + t] = d.xn[(size_t)i * K + t];
    }
    for (int idx = threadIdx.x; idx < (N - 1) * NU; idx += blockDim.x) {
      const int i = idx / (N - 1), k = idx % (N - 1);
      const size_t t = (size_t)b * N + k;
      d.u[(size_t)i * K + t] = d.un[(size_t)i * K + t];
    }
  }
}

// -----------------------------------------------------------------------------------------------------------------
// k_linesearch: the whole backtracking line search of one SQP iteration (SQP :606-744) plus check_for_exit_or_error
// (:463-481), one block per active instance, one thread per knot point.  Each trial: trial point x - alpha dz (kept
// in registers), forward dynamics, cost / penalty / defect / directional-derivative terms per knot in parallel, then
// thread 0 sums them in the reference's sequential order and decides accept / halve alpha / fail.  Replaces the
// 2 x max_trials launches of k_fd<TRIAL> + k_merit and the k_sqp_ctrl launch.
// smem: (6 + NX) * N scalars.
// -----------------------------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ void outer_update(const Dev<T>& d, const Opts<T>& o, int b, T* sm, int fused_restart);   // defined below

template <typename T, bool MS = false>
__global__ void __launch_bounds__(128) k_linesearch(Dev<T> d, Opts<T> o, int fuse_outer, int fuse_recover) {
  if ((int)blockIdx.x >= *d.n_act) return;
  const int b = d.act[blockIdx.x];
  const int N = d.N;
  const size_t K = d.K;
  extern __shared__ unsigned char smem_raw[];
  T* sm = reinterpret_cast<T*>(smem_raw);
  // the step dz_k of the structured path is computed here (what k_recover_diag does in its own launch): each thread produces
  // the knots it evaluates below, so no barrier is needed
  if (fuse_recover)
    for (int k = threadIdx.x; k < N; k += blockDim.x) recover_diag_knot(d, b, k);
  T* s_cost = sm; T* s_soft = sm + N; T* s_c = sm + 2 * N; T* s_D = sm + 3 * N; T* s_Ds = sm + 4 * N; T* s_xn = sm + 5 * N;   // [N][NX]
  T* s_h = s_xn + N * NX;     // [N] violated hard bounds (ACTIVE_SET)
  __shared__ int s_state;     // 0: accepted, 1: try a smaller alpha, 2: failed
  __shared__ int s_exit;      // the SQP loop of this instance exited
  T alpha = T(1);
  int ls = 0;
  for (;;) {
    for (int k = threadIdx.x; k < N; k += blockDim.x) {
      const size_t t = (size_t)b * N + k;
      const bool terminal = (k == N - 1);
      T z[NM], dzk[NM], xg[NX];
      load_xu(d.x, d.u, K, t, terminal, z, z + NX);
      for (int i = 0; i < NM; ++i) dzk[i] = d.dz[(size_t)i * K + t];
      for (int i = 0; i < NX; ++i) { z[i] = z[i] - alpha * dzk[i]; d.xn[(size_t)i * K + t] = z[i]; xg[i] = d.xg[(size_t)i * d.B + b]; }
      if (!terminal)
        for (int i = 0; i < NU; ++i) { z[NX + i] = z[NX + i] - alpha * dzk[NX + i]; d.un[(size_t)i * K + t] = z[NX + i]; }
      const int M = terminal ? NX : NM;
      if (d.diag_mode) {        // diagonal weights: value and g . dz in one pass (bit-identical to the general functions)
        T cv, cd;
        cost_value_dir_diag(d.cost, z, z + NX, xg, k, terminal, dzk, &cv, &cd);
        s_cost[k] = cv; s_D[k] = cd;
      } else {
        s_cost[k] = cost_value(d.cost, z, z + NX, xg, k, terminal);
        T g[NM];
        cost_grad_hess<T, false>(d.cost, z, z + NX, xg, k, terminal, g, (T*)nullptr);
        T acc = T(0);
        for (int i = 0; i < M; ++i) acc += g[i] * dzk[i];
        s_D[k] = acc;
      }
      T sv = T(0), accs = T(0);
      if (d.lim.any) soft_value_dir(d.lim, z, d.mu + t, d.lam + t, K, terminal, dzk, &sv, &accs);
      s_soft[k] = sv;
      s_Ds[k] = accs;
      if (d.hard.any) s_h[k] = hard_violation(d.hard, z, terminal);
      if (k == 0) {
        T cc = T(0);
        for (int i = 0; i < NX; ++i) cc += fabs(z[i] - d.xs[(size_t)i * d.B + b]);
        s_c[0] = cc;
      }
      // dynamics last: dz, xg and the multipliers are dead by now, which keeps the recursion's temporaries in registers
      if (!terminal) {
        T qdd[NJ], Minv[NJ * NJ], v[NJ][6], a[NJ][6], f[NJ][6], xnext[NX];
        if constexpr (MS) integrator_multi_value(d.integrator, z, z + NX, d.gravity, d.dt, xnext);
        else {
          forward_dynamics<T, false>(z, z + NJ, z + NX, d.gravity, qdd, Minv, v, a, f);
          integrate(d.integrator, z, qdd, d.dt, xnext);
        }
        for (int i = 0; i < NX; ++i) s_xn[k * NX + i] = xnext[i];
      }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < N; k += blockDim.x) {         // defect of knot k against the step from knot k-1
      if (k == 0) continue;
      const size_t t = (size_t)b * N + k;
      T cc = T(0);
      for (int i = 0; i < NX; ++i) cc += fabs(d.xn[(size_t)i * K + t] - s_xn[(k - 1) * NX + i]);
      s_c[k] = cc;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      // J (costs, then penalties), c and D each summed in the reference's sequential order; the three dependent chains are
      // interleaved in one loop so that their add latencies overlap
      T Jn = T(0), cn = T(0), D = T(0);
      const bool lim_any = d.lim.any != 0;
      for (int k = 0; k < N; ++k) {
        Jn += s_cost[k];
        cn += s_c[k];
        D += s_D[k];
        if (lim_any) D += s_Ds[k];
      }
      if (lim_any) {
        for (int k = 0; k < N; ++k) Jn += s_soft[k];
      }
      if (d.hard.any)
        for (int k = 0; k < N; ++k) cn += s_h[k];
      const T mu = o.merit_mu;
      const T merit_new = Jn + mu * cn;
      const T delta_J = d.J[b] - Jn;
      const T delta_merit = d.merit[b] - merit_new;
      const T expected = alpha * (D - mu * cn);
      const T ratio = delta_merit / expected;
      d.deltaJ[b] = delta_J;
      d.tot_trials[b] += 1;
      if (delta_merit >= T(0) && ratio >= o.er_min && ratio <= o.er_max) {
        s_state = 0;
        d.J[b] = Jn; d.c[b] = cn; d.merit[b] = merit_new;
        const T drho = fmin(d.drho[b] / o.rho_factor, T(1) / o.rho_factor);     // reduce_regularization (:457-461)
        d.drho[b] = drho; d.rho[b] = fmax(d.rho[b] * drho, o.rho_min);
        trace_row(d, b, ls, alpha, D, ratio, d.pcg_iters[b], 1);
      } else if (alpha > o.alpha_min) {
        s_state = 1;
      } else {
        s_state = 2;
        trace_row(d, b, ls, alpha, D, ratio, d.pcg_iters[b], 0);
      }
    }
    __syncthreads();
    const int st = s_state;
    if (st == 0) {
      for (int idx = threadIdx.x; idx < N * NX; idx += blockDim.x) {
        const int i = idx / N, k = idx % N;
        const size_t t = (size_t)b * N + k;
        d.x[(size_t)i * K + t] = d.xn[(size_t)i * K + t];
      }
      for (int idx = threadIdx.x; idx < (N - 1) * NU; idx += blockDim.x) {
        const int i = idx / (N - 1), k = idx % (N - 1);
        const size_t t = (size_t)b * N + k;
        d.u[(size_t)i * K + t] = d.un[(size_t)i * K + t];
      }
    }
    if (st != 1) {
      if (threadIdx.x == 0) {           // check_for_exit_or_error (:463-481)
        bool exit_flag = false;
        if (st == 2) {
          const T drho = fmax(d.drho[b] * o.rho_factor, o.rho_factor);
          const T rho = fmax(d.rho[b] * drho, o.rho_min);
          d.drho[b] = drho; d.rho[b] = rho;
          if (rho > o.rho_max) { d.exit_sqp[b] = 2; exit_flag = true; }
        } else if (d.deltaJ[b] < o.tol_sqp) {
          d.exit_sqp[b] = 1; exit_flag = true;
        }
        if (d.sqp_iter[b] == o.max_iter_sqp - 1) { d.exit_sqp[b] = 3; exit_flag = true; }
        else d.sqp_iter[b] += 1;
        if (exit_flag) d.phase[b] = PH_OUTER;
        d.ls_iter[b] = ls;
        d.alpha[b] = alpha;
        d.dyn_ok[b] = (st == 2) ? 1 : 0;      // a failed search leaves x, u (hence A, B, x+) untouched; only rho changes
        s_exit = exit_flag ? 1 : 0;
      }
      __syncthreads();
      // the SQP loop of this instance exited: soft-constraint check / update and the restart of the next outer iteration right here
      // (what a separate k_outer launch over all active instances did for the few that exit in a pass)
      if (fuse_outer && s_exit) outer_update(d, o, b, sm, 1);
      return;
    }
    alpha *= o.alpha_factor;
    ls += 1;
    __syncthreads();      // s_state / shared terms are rewritten by the next trial
  }
}

// -----------------------------------------------------------------------------------------------------------------
// k_linesearch_par: the same line search with ALL trial step lengths of an instance evaluated at once, one block per
// (instance, trial) -- for passes with few active instances, where the sequential kernel costs its latency (~12 us per trial,
// ~7 trials per search on the benchmark workload) while most SMs idle.  Trial t evaluates alpha_t = alpha_factor^t exactly like
// iteration t of the sequential loop and stores (J, c, D); the LAST block of an instance to finish (atomic counter) walks the
// trials in order, takes the sequential loop's decision at the first trial that accepts or fails -- later trials are discarded, so
// counts and results are bit-identical (tests/test_gpu_variants.py) -- writes the accepted point x - alpha dz and runs the exit
// tests / outer update.  Work grows by max_trials / (trials used); the host selects it only below a small active count.
// smem: (6 + 2 NX) * N scalars.   scratch: trial_out [B][max_trials][3], done [B] (zero between passes).
// -----------------------------------------------------------------------------------------------------------------
template <typename T, bool MS = false>
__global__ void __launch_bounds__(128) k_linesearch_par(Dev<T> d, Opts<T> o, int fuse_outer, int fuse_recover, int max_trials, T* trial_out, int* done) {
  const int slot = (int)blockIdx.x / max_trials, trial = (int)blockIdx.x % max_trials;
  if (slot >= *d.n_act) return;
  const int b = d.act[slot];
  const int N = d.N;
  const size_t K = d.K;
  extern __shared__ unsigned char smem_raw[];
  T* sm = reinterpret_cast<T*>(smem_raw);
  if (fuse_recover)        // every trial block of the instance computes the same dz (identical values; the write is idempotent)
    for (int k = threadIdx.x; k < N; k += blockDim.x) recover_diag_knot(d, b, k);
  T* s_cost = sm; T* s_soft = sm + N; T* s_c = sm + 2 * N; T* s_D = sm + 3 * N; T* s_Ds = sm + 4 * N; T* s_xn = sm + 5 * N;   // [N][NX]
  T* s_h = s_xn + N * NX;     // [N]
  T* s_x = s_h + N;           // [N][NX] trial states
  __shared__ int s_last, s_state, s_exit, s_ls;
  __shared__ T s_alpha;
  T alpha = T(1);
  for (int i = 0; i < trial; ++i) alpha *= o.alpha_factor;        // the sequential loop's alpha after `trial` halvings, bit for bit
  for (int k = threadIdx.x; k < N; k += blockDim.x) {
    const size_t t = (size_t)b * N + k;
    const bool terminal = (k == N - 1);
    T z[NM], dzk[NM], xg[NX];
    load_xu(d.x, d.u, K, t, terminal, z, z + NX);
    for (int i = 0; i < NM; ++i) dzk[i] = d.dz[(size_t)i * K + t];
    for (int i = 0; i < NX; ++i) { z[i] = z[i] - alpha * dzk[i]; s_x[k * NX + i] = z[i]; xg[i] = d.xg[(size_t)i * d.B + b]; }
    if (!terminal)
      for (int i = 0; i < NU; ++i) z[NX + i] = z[NX + i] - alpha * dzk[NX + i];
    const int M = terminal ? NX : NM;
    if (d.diag_mode) {
      T cv, cd;
      cost_value_dir_diag(d.cost, z, z + NX, xg, k, terminal, dzk, &cv, &cd);
      s_cost[k] = cv; s_D[k] = cd;
    } else {
      s_cost[k] = cost_value(d.cost, z, z + NX, xg, k, terminal);
      T g[NM];
      cost_grad_hess<T, false>(d.cost, z, z + NX, xg, k, terminal, g, (T*)nullptr);
      T acc = T(0);
      for (int i = 0; i < M; ++i) acc += g[i] * dzk[i];
      s_D[k] = acc;
    }
    T sv = T(0), accs = T(0);
    if (d.lim.any) soft_value_dir(d.lim, z, d.mu + t, d.lam + t, K, terminal, dzk, &sv, &accs);
    s_soft[k] = sv;
    s_Ds[k] = accs;
    if (d.hard.any) s_h[k] = hard_violation(d.hard, z, terminal);
    if (k == 0) {
      T cc = T(0);
      for (int i = 0; i < NX; ++i) cc += fabs(z[i] - d.xs[(size_t)i * d.B + b]);
      s_c[0] = cc;
    }
    if (!terminal) {
      T qdd[NJ], Minv[NJ * NJ], v[NJ][6], a[NJ][6], f[NJ][6], xnext[NX];
      if constexpr (MS) integrator_multi_value(d.integrator, z, z + NX, d.gravity, d.dt, xnext);
      else {
        forward_dynamics<T, false>(z, z + NJ, z + NX, d.gravity, qdd, Minv, v, a, f);
        integrate(d.integrator, z, qdd, d.dt, xnext);
      }
      for (int i = 0; i < NX; ++i) s_xn[k * NX + i] = xnext[i];
    }
  }
  __syncthreads();
  for (int k = threadIdx.x; k < N; k += blockDim.x) {
    if (k == 0) continue;
    T cc = T(0);
    for (int i = 0; i < NX; ++i) cc += fabs(s_x[k * NX + i] - s_xn[(k - 1) * NX + i]);
    s_c[k] = cc;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    T Jn = T(0), cn = T(0), D = T(0);
    const bool lim_any = d.lim.any != 0;
    for (int k = 0; k < N; ++k) {
      Jn += s_cost[k];
      cn += s_c[k];
      D += s_D[k];
      if (lim_any) D += s_Ds[k];
    }
    if (lim_any) {
      for (int k = 0; k < N; ++k) Jn += s_soft[k];
    }
    if (d.hard.any)
      for (int k = 0; k < N; ++k) cn += s_h[k];
    T* out = trial_out + ((size_t)b * max_trials + trial) * 3;
    out[0] = Jn; out[1] = cn; out[2] = D;
    __threadfence();
    const int prev = atomicAdd(done + b, 1);
    s_last = (prev == max_trials - 1) ? 1 : 0;
    if (s_last) { done[b] = 0; __threadfence(); }
  }
  __syncthreads();
  if (!s_last) return;
  // ---- last block of this instance: the sequential loop's decisions over the stored trials
  if (threadIdx.x == 0) {
    T al = T(1);
    int ls = 0, state = 2;
    const T mu = o.merit_mu;
    for (int t = 0; t < max_trials; ++t) {
      const volatile T* in = trial_out + ((size_t)b * max_trials + t) * 3;
      const T Jn = in[0], cn = in[1], D = in[2];
      const T merit_new = Jn + mu * cn;
      const T delta_J = d.J[b] - Jn;
      const T delta_merit = d.merit[b] - merit_new;
      const T expected = al * (D - mu * cn);
      const T ratio = delta_merit / expected;
      d.deltaJ[b] = delta_J;
      d.tot_trials[b] += 1;
      if (delta_merit >= T(0) && ratio >= o.er_min && ratio <= o.er_max) {
        state = 0;
        d.J[b] = Jn; d.c[b] = cn; d.merit[b] = merit_new;
        const T drho = fmin(d.drho[b] / o.rho_factor, T(1) / o.rho_factor);
        d.drho[b] = drho; d.rho[b] = fmax(d.rho[b] * drho, o.rho_min);
        trace_row(d, b, ls, al, D, ratio, d.pcg_iters[b], 1);
        break;
      } else if (al > o.alpha_min) {
        al *= o.alpha_factor;
        ls += 1;
      } else {
        state = 2;
        trace_row(d, b, ls, al, D, ratio, d.pcg_iters[b], 0);
        break;
      }
    }
    s_state = state; s_alpha = al; s_ls = ls;
  }
  __syncthreads();
  const int st = s_state;
  const T al = s_alpha;
  // the deciding trial's point (what the sequential kernel leaves in xn / un), and on acceptance the new iterate
  for (int k = threadIdx.x; k < N; k += blockDim.x) {
    const size_t t = (size_t)b * N + k;
    const bool terminal = (k == N - 1);
    for (int i = 0; i < NX; ++i) {
      const T v = d.x[(size_t)i * K + t] - al * d.dz[(size_t)i * K + t];
      d.xn[(size_t)i * K + t] = v;
      if (st == 0) d.x[(size_t)i * K + t] = v;
    }
    if (!terminal)
      for (int i = 0; i < NU; ++i) {
        const T v = d.u[(size_t)i * K + t] - al * d.dz[(size_t)(NX + i) * K + t];
        d.un[(size_t)i * K + t] = v;
        if (st == 0) d.u[(size_t)i * K + t] = v;
      }
  }
  if (threadIdx.x == 0) {           // check_for_exit_or_error (:463-481)
    bool exit_flag = false;
    if (st == 2) {
      const T drho = fmax(d.drho[b] * o.rho_factor, o.rho_factor);
      const T rho = fmax(d.rho[b] * drho, o.rho_min);
      d.drho[b] = drho; d.rho[b] = rho;
      if (rho > o.rho_max) { d.exit_sqp[b] = 2; exit_flag = true; }
    } else if (d.deltaJ[b] < o.tol_sqp) {
      d.exit_sqp[b] = 1; exit_flag = true;
    }
    if (d.sqp_iter[b] == o.max_iter_sqp - 1) { d.exit_sqp[b] = 3; exit_flag = true; }
    else d.sqp_iter[b] += 1;
    if (exit_flag) d.phase[b] = PH_OUTER;
    d.ls_iter[b] = s_ls;
    d.alpha[b] = al;
    d.dyn_ok[b] = (st == 2) ? 1 : 0;
    s_exit = exit_flag ? 1 : 0;
  }
  __syncthreads();
  if (fuse_outer && s_exit) outer_update(d, o, b, sm, 1);
}

// stage entry point: J, c, D of the trial point for alpha (no decision)
template <typename T>
__global__ void k_merit_only(Dev<T> d, T* J, T* c, T* D) {
  const int b = blockIdx.x;
  extern __shared__ unsigned char smem_raw[];
  T* sm = reinterpret_cast<T*>(smem_raw);
  __shared__ T Jn, cn, Dn;
  merit_terms<T, true, true, true>(d, b, sm, &Jn, &cn, &Dn);
  if (threadIdx.x == 0) { J[b] = Jn; c[b] = cn; D[b] = Dn; }
}

// -----------------------------------------------------------------------------------------------------------------
// k_sqp_ctrl: check_for_exit_or_error (:463-481) per active instance
// -----------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void k_sqp_ctrl(Dev<T> d, Opts<T> o) {
  const int n = *d.n_act;
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  const int b = d.act[s];
  bool exit_flag = false;
  if (d.err[b]) {
    T drho = fmax(d.drho[b] * o.rho_factor, o.rho_factor);
    T rho = fmax(d.rho[b] * drho, o.rho_min);
    d.drho[b] = drho; d.rho[b] = rho;
    if (rho > o.rho_max) { d.exit_sqp[b] = 2; exit_flag = true; }
  } else if (d.deltaJ[b] < o.tol_sqp) {
    d.exit_sqp[b] = 1; exit_flag = true;
  }
  if (d.sqp_iter[b] == o.max_iter_sqp - 1) { d.exit_sqp[b] = 3; exit_flag = true; }
  else d.sqp_iter[b] += 1;
  if (exit_flag) d.phase[b] = PH_OUTER;
}

// -----------------------------------------------------------------------------------------------------------------
// k_outer: check_and_update_soft_constraints (:483-508) for instances whose SQP loop exited; BoxConstraint
// max_soft_constraint_value (TrajoptConstraint.py:131-136) and update_soft_constraint_constants (:138-166).
// Instances that continue re-enter the SQP loop through k_outer_begin (the host relaunches it on the `restart` list).
// -----------------------------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ void outer_update(const Dev<T>& d, const Opts<T>& o, int b, T* sm, int fused_restart) {
  int* restart_list = d.restart_list;
  int* restart_count = d.n_restart;
  const int N = d.N;
  const size_t K = d.K;
  __shared__ T s_max[3];
  __shared__ int s_flag;
  __shared__ int s_continue;
  if (threadIdx.x == 0) { s_max[0] = s_max[1] = s_max[2] = T(0); s_flag = 1; s_continue = 0; }
  __syncthreads();
  T max_c = T(0);
  if (d.lim.any) {
    // per limit type: max over knots of | min over the type's 2*cs values |;  sm: [3][N]
    for (int k = threadIdx.x; k < N; k += blockDim.x) {
      const size_t t = (size_t)b * N + k;
      const bool terminal = (k == N - 1);
      T z[NM];
      load_xu(d.x, d.u, K, t, terminal, z, z + NX);
      const int lo[3] = {0, NJ, NX}, hi[3] = {NJ, NX, NM};
      for (int ty = 0; ty < 3; ++ty) {
        T mn = T(0);
        bool has = false;
        if (!(ty == 2 && terminal)) {
          for (int i = lo[ty]; i < hi[ty]; ++i) {
            if (d.lim.mode[i] == LIM_NONE) continue;
            T vlo = z[i] - d.lim.lb[i], vhi = d.lim.ub[i] - z[i];
            T m2 = fmin(vlo, vhi);
            mn = has ? fmin(mn, m2) : m2;
            has = true;
          }
        }
        sm[ty * N + k] = has ? fabs(mn) : T(-1);
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int ty = 0; ty < 3; ++ty)
        for (int k = 0; k < N; ++k) max_c = fmax(max_c, sm[ty * N + k]);
    }
  }
  if (threadIdx.x == 0) {
    bool exit_flag = false;
    if (max_c < o.tol_soft) { d.exit_soft[b] = 1; exit_flag = true; }
    if (d.outer_iter[b] == o.max_iter_soft - 1) { d.exit_soft[b] = 2; exit_flag = true; }
    else d.outer_iter[b] += 1;
    s_continue = exit_flag ? 0 : 1;
  }
  __syncthreads();
  if (!s_continue) {
    if (threadIdx.x == 0) d.phase[b] = PH_DONE;
    return;
  }
  // update mu / lambda / phi  (all limit types, every knot, both sides)
  int changed = 0;
  for (int k = threadIdx.x; k < N; k += blockDim.x) {
    const size_t t = (size_t)b * N + k;
    const bool terminal = (k == N - 1);
    T z[NM];
    load_xu(d.x, d.u, K, t, terminal, z, z + NX);
    const int lim = terminal ? NX : NM;
    for (int i = 0; i < lim; ++i) {
      if (d.lim.mode[i] == LIM_NONE) continue;
      const int ty = i < NJ ? 0 : (i < NX ? 1 : 2);
      for (int side = 0; side < 2; ++side) {
        const T v = side == 0 ? z[i] - d.lim.lb[i] : d.lim.ub[i] - z[i];
        if (!(v < T(0))) continue;
        const size_t idx = (size_t)(side * NM + i) * K + t;
        if (!(fabs(v) < d.phi[idx])) {
          const T cur = d.mu[idx];
          if (cur < d.mu_max[ty]) { changed = 1; d.mu[idx] = fmin(d.mu_max[ty], cur * d.mu_factor[ty]); }
        } else {
          changed = 1;
          d.lam[idx] += d.mu[idx] * v;
          d.phi[idx] /= d.phi_factor[ty];
        }
      }
    }
  }
  if (changed) atomicAnd(&s_flag, 0);
  __syncthreads();
  if (s_flag) {
    if (threadIdx.x == 0) { d.exit_soft[b] = 3; d.phase[b] = PH_DONE; }
    return;
  }
  if (!fused_restart) {
    if (threadIdx.x == 0) {
      const int pos = atomicAdd(restart_count, 1);
      restart_list[pos] = b;
    }
    return;
  }
  // next outer iteration (SQP :535-569): J with the updated penalties, rho, merit, seed trace row; c is unchanged
  __threadfence_block();
  __syncthreads();
  {
    __shared__ T Jb, cb, Db;
    merit_terms<T, false, false, false>(d, b, sm, &Jb, &cb, &Db);
    if (threadIdx.x == 0) {
      d.J[b] = Jb;
      d.rho[b] = o.rho_init;
      d.drho[b] = T(1);
      d.merit[b] = Jb + o.merit_mu * d.c[b];
      d.sqp_iter[b] = 0;
      d.trace_rows[b] = 0;
      d.phase[b] = PH_SQP;
      trace_row(d, b, 0, T(1), T(0), T(0), 0, 0);
    }
  }
}

// standalone launch of outer_update (legacy line-search path); smem: max(3, 5) * N scalars
template <typename T>
__global__ void k_outer(Dev<T> d, Opts<T> o, int fused_restart) {
  if ((int)blockIdx.x >= *d.n_act) return;
  const int b = d.act[blockIdx.x];
  if (d.phase[b] != PH_OUTER) return;
  extern __shared__ unsigned char smem_raw[];
  outer_update(d, o, b, reinterpret_cast<T*>(smem_raw), fused_restart);
}

// rebuild the active list (order-preserving, single block of 1024 threads: per-thread chunk counts + block scan)
template <typename T>
__global__ void __launch_bounds__(1024) k_compact(Dev<T> d, int* scratch, int* pass_trace = nullptr, int pass = 0) {
  __shared__ int s_cnt[1024];
  const int n = *d.n_act;
  const int chunk = (n + blockDim.x - 1) / blockDim.x;
  const int lo = min(n, (int)threadIdx.x * chunk), hi = min(n, lo + chunk);
  int cnt = 0;
  for (int s = lo; s < hi; ++s) cnt += (d.phase[d.act[s]] != PH_DONE) ? 1 : 0;
  s_cnt[threadIdx.x] = cnt;
  __syncthreads();
  for (int off = 1; off < (int)blockDim.x; off <<= 1) {       // inclusive Hillis-Steele scan
    int v = (threadIdx.x >= (unsigned)off) ? s_cnt[threadIdx.x - off] : 0;
    __syncthreads();
    s_cnt[threadIdx.x] += v;
    __syncthreads();
  }
  int pos = s_cnt[threadIdx.x] - cnt;
  for (int s = lo; s < hi; ++s) {
    const int b = d.act[s];
    if (d.phase[b] != PH_DONE) scratch[pos++] = b;
  }
  const int total = s_cnt[blockDim.x - 1];
  __syncthreads();
  for (int s = threadIdx.x; s < total; s += blockDim.x) d.act[s] = scratch[s];
  __syncthreads();
  if (threadIdx.x == 0) { *d.n_act = total; *d.n_restart = 0; if (pass_trace) pass_trace[pass] = total; }
}

// -----------------------------------------------------------------------------------------------------------------
// layout conversion between the ABI (reference) layouts and the SoA workspace
// -----------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void k_pack_traj(Dev<T> d, const double* x, const double* u) {   // x [B][NX][N], u [B][NU][N-1]
  const size_t gt = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gt >= d.K) return;
  const int b = (int)(gt / d.N), k = (int)(gt % d.N);
  for (int i = 0; i < NX; ++i) d.x[(size_t)i * d.K + gt] = (T)x[((size_t)b * NX + i) * d.N + k];
  for (int i = 0; i < NU; ++i) d.u[(size_t)i * d.K + gt] = (k < d.N - 1) ? (T)u[((size_t)b * NU + i) * (d.N - 1) + k] : T(0);
  if (k == 0)
    for (int i = 0; i < NX; ++i) d.xs[(size_t)i * d.B + b] = (T)x[((size_t)b * NX + i) * d.N];
}
template <typename T>
__global__ void k_unpack_traj(Dev<T> d, double* x, double* u) {
  const size_t gt = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gt >= d.K) return;
  const int b = (int)(gt / d.N), k = (int)(gt % d.N);
  for (int i = 0; i < NX; ++i) x[((size_t)b * NX + i) * d.N + k] = (double)d.x[(size_t)i * d.K + gt];
  if (k < d.N - 1)
    for (int i = 0; i < NU; ++i) u[((size_t)b * NU + i) * (d.N - 1) + k] = (double)d.u[(size_t)i * d.K + gt];
}
template <typename T>
__global__ void k_pack_goals(Dev<T> d, const double* xg) {   // [B][NX]
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= d.B) return;
  for (int i = 0; i < NX; ++i) d.xg[(size_t)i * d.B + b] = (T)xg[(size_t)b * NX + i];
}
// generic fetch of a per-knot SoA array a[E][K] into knot-major doubles out[K][E]
template <typename T>
__global__ void k_fetch_soa(const T* a, size_t K, int E, double* out) {
  const size_t gt = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gt >= K) return;
  for (int e = 0; e < E; ++e) out[gt * E + e] = (double)a[(size_t)e * K + gt];
}

// per-knot cost value / gradient / Hessian at (x, u) -> knot-major doubles (what = 0, 1, 2)
template <typename T>
__global__ void k_cost_eval(Dev<T> d, int what, double* out) {
  const size_t gt = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gt >= d.K) return;
  const int b = (int)(gt / d.N), k = (int)(gt % d.N);
  const bool terminal = (k == d.N - 1);
  T z[NM], xg[NX];
  load_xu(d.x, d.u, d.K, gt, terminal, z, z + NX);
  for (int i = 0; i < NX; ++i) xg[i] = d.xg[(size_t)i * d.B + b];
  if (what == 0) { out[gt] = (double)cost_value(d.cost, z, z + NX, xg, k, terminal); return; }
  if (what == 3) {      // state error of the cost's state map (UrdfCost.delta_x, TrajoptCost.py:425-435)
    T e[NX], Jt[NX * NX]; bool hj;
    const int ne = cost_state(d.cost, z, xg, e, Jt, &hj);
    for (int i = 0; i < NX; ++i) out[gt * NX + i] = i < ne ? (double)e[i] : 0.0;
    return;
  }
  if (what == 8) {      // Jacobian of the cost's state map (UrdfCost: jacobian_tot_state, TrajoptCost.py:439), ne x nx row-major in an nx*nx slot
    T e[NX], Jt[NX * NX]; bool hj;
    const int ne = cost_state(d.cost, z, xg, e, Jt, &hj);
    for (int i = 0; i < NX * NX; ++i) out[gt * NX * NX + i] = 0.0;
    for (int i = 0; i < ne; ++i)
      for (int j = 0; j < NX; ++j) out[gt * NX * NX + i * NX + j] = hj ? (double)Jt[i * NX + j] : (i == j ? 1.0 : 0.0);
    return;
  }
  if (what == 6 || what == 7) {      // soft box limits at (x_k, u_k) with the current multipliers: value, resp. summed gradient gck (NM)
    T gck[NM];
    for (int i = 0; i < NM; ++i) gck[i] = T(0);
    T val = T(0);
    if (d.lim.any) {
      val = soft_value(d.lim, z, d.mu + gt, d.lam + gt, d.K, terminal);
      soft_grad(d.lim, z, d.mu + gt, d.lam + gt, d.K, terminal, gck);
    }
    if (what == 6) out[gt] = (double)val;
    else for (int i = 0; i < NM; ++i) out[gt * NM + i] = (double)gck[i];
    return;
  }
  if (what == 5) {      // [A_k B_k] of the integrator at (x_k, u_k) from the stored forward-dynamics gradient (row N-1 unused: zeros)
    T AB[NX * NM];
    if (terminal) { for (int i = 0; i < NX * NM; ++i) AB[i] = T(0); }
    else load_AB(d, gt, AB);
    for (int i = 0; i < NX * NM; ++i) out[gt * NX * NM + i] = terminal ? 0.0 : (double)AB[i];
    return;
  }
  T g[NM], H[NM * NM];
  cost_grad_hess<T, true>(d.cost, z, z + NX, xg, k, terminal, g, H);
  if (what == 4 && d.lim.any) {      // KKT Hessian block G_k = cost Hessian + gck gck^T  (TrajoptMPCReference.py:220-224), without rho
    T gck[NM];
    soft_grad(d.lim, z, d.mu + gt, d.lam + gt, d.K, terminal, gck);
    for (int i = 0; i < NM; ++i)
      for (int j = 0; j < NM; ++j) H[i * NM + j] += gck[i] * gck[j];
  }
  if (what == 1) { for (int i = 0; i < NM; ++i) out[gt * NM + i] = (double)g[i]; }
  else { for (int i = 0; i < NM * NM; ++i) out[gt * NM * NM + i] = (double)H[i]; }
}
// plant-level terms of one knot at (x_k, u_k), what the reference's URDFPlant callbacks append to saved_c / saved_qdd / saved_Minv /
// saved_dc_du (TrajoptPlant.py:283-323): out[knot] = [c (n) | qdd (n) | Minv (n x n) | dc/d(q, qd) (n x 2n)], row-major doubles
constexpr int PLANT_TERMS = 2 * NJ + NJ * NJ + 2 * NJ * NJ;
template <typename T>
__global__ void __launch_bounds__(64) k_plant_eval(Dev<T> d, double* out) {
  const size_t gt = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gt >= d.K) return;
  const int k = (int)(gt % d.N);
  double* o = out + gt * PLANT_TERMS;
  if (k == d.N - 1) { for (int i = 0; i < PLANT_TERMS; ++i) o[i] = 0.0; return; }
  T z[NM];
  load_xu(d.x, d.u, d.K, gt, false, z, z + NX);
  T xe[NJ][NXE], c[NJ], qdd[NJ], Minv[NJ * NJ], v[NJ][6], a[NJ][6], f[NJ][6];
  xset_all(z, xe);
  rnea<T, false>(xe, z + NJ, nullptr, d.gravity, v, a, f, c);
  for (int i = 0; i < NJ; ++i) o[i] = (double)c[i];
  forward_dynamics<T, true>(z, z + NJ, z + NX, d.gravity, qdd, Minv, v, a, f);      // leaves v, a, f of rnea(q, qd, qdd)
  for (int i = 0; i < NJ; ++i) o[NJ + i] = (double)qdd[i];
  for (int i = 0; i < NJ * NJ; ++i) o[2 * NJ + i] = (double)Minv[i];
  for (int col = 0; col < 2 * NJ; ++col) {
    T dq[NJ], dc[NJ];
    fd_grad_column(z, z + NJ, v, a, f, Minv, d.gravity, col % NJ, col >= NJ, dq, dc);
    for (int i = 0; i < NJ; ++i) o[2 * NJ + NJ * NJ + i * 2 * NJ + col] = (double)dc[i];
  }
}
// standalone PCG entry (PCG(A, b, block_size, Nblocks).solve()): upload of a block-tridiagonal system in knot-major doubles
template <typename T>
__global__ void k_set_block_system(Dev<T> d, const double* Sd, const double* So, const double* gam) {
  const size_t gt = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gt >= d.K) return;
  for (int e = 0; e < NX * NX; ++e) { d.Sd[(size_t)e * d.K + gt] = (T)Sd[gt * NX * NX + e]; d.So[(size_t)e * d.K + gt] = (T)So[gt * NX * NX + e]; }
  for (int i = 0; i < NX; ++i) d.gam[(size_t)i * d.K + gt] = (T)gam[gt * NX + i];
}
template <typename T>
__global__ void k_fill(T* p, size_t n, T v) {
  const size_t gt = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gt < n) p[gt] = v;
}
// multipliers: ABI layout [B][2m][N] <-> SoA [2m][K]
template <typename T>
__global__ void k_pack_mult(Dev<T> d, const double* src, T* dst) {
  const size_t gt = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gt >= d.K) return;
  const int b = (int)(gt / d.N), k = (int)(gt % d.N);
  for (int e = 0; e < 2 * NM; ++e) dst[(size_t)e * d.K + gt] = (T)src[((size_t)b * 2 * NM + e) * d.N + k];
}
template <typename T>
__global__ void k_unpack_mult(Dev<T> d, const T* src, double* dst) {
  const size_t gt = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gt >= d.K) return;
  const int b = (int)(gt / d.N), k = (int)(gt % d.N);
  for (int e = 0; e < 2 * NM; ++e) dst[((size_t)b * 2 * NM + e) * d.N + k] = (double)src[(size_t)e * d.K + gt];
}
template <typename T>
__global__ void k_init_mult(Dev<T> d) {
  const size_t gt = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gt >= d.K) return;
  for (int e = 0; e < 2 * NM; ++e) {
    const int i = e % NM;
    const int ty = i < NJ ? 0 : (i < NX ? 1 : 2);
    d.mu[(size_t)e * d.K + gt] = d.mu_init[ty];
    d.lam[(size_t)e * d.K + gt] = T(0);
    d.phi[(size_t)e * d.K + gt] = d.phi_init[ty];
  }
}
template <typename T>
__global__ void k_init_state(Dev<T> d) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b == 0) *d.n_act = d.B;
  if (b >= d.B) return;
  d.act[b] = b;
  d.dyn_ok[b] = 0;
  d.outer_iter[b] = 0; d.sqp_iter[b] = 0; d.exit_sqp[b] = 0; d.exit_soft[b] = 0; d.phase[b] = PH_SQP; d.err[b] = 0;
  d.pcg_iters[b] = 0; d.tot_qp[b] = 0; d.tot_pcg[b] = 0; d.tot_trials[b] = 0; d.trace_rows[b] = 0; d.ls_iter[b] = 0;
  d.alpha[b] = T(1); d.deltaJ[b] = T(0); d.c[b] = T(0); d.J[b] = T(0); d.merit[b] = T(0); d.rho[b] = T(0); d.drho[b] = T(1);
}

}  // namespace b2t

namespace b2t {
// -----------------------------------------------------------------------------------------------------------------
// k_mpc_shift: receding-horizon shift by one knot: the applied control is u_0, the next initial state is
// either measured (x_next given) or simulated with the plant's own integrator; x, u move one knot to the left (last knot
// repeated) as the warm start of the next solve; the soft-constraint multipliers move with them
// (TrajoptConstraint.shift_soft_constraint_constants, TrajoptConstraint.py:168-176,380-387; the last column is re-initialised --
// the reference's slice `[:, shift_steps:] = init` re-initialises every column but the first).
// One block per instance, one thread per knot.  out_x0 / out_u0 (may be null): the state before the shift and the applied control.
// -----------------------------------------------------------------------------------------------------------------
template <typename T, bool MS = false>
__global__ void k_mpc_shift(Dev<T> d, const double* x_next, double* out_x0, double* out_u0, double* out_xnext) {
  const int b = blockIdx.x;
  const int N = d.N;
  const size_t K = d.K;
  const size_t t0 = (size_t)b * N;
  __shared__ T s_xn[NX];
  if (threadIdx.x == 0) {
    T x[NX], u[NU];
    for (int i = 0; i < NX; ++i) x[i] = d.x[(size_t)i * K + t0];
    for (int i = 0; i < NU; ++i) u[i] = d.u[(size_t)i * K + t0];
    if (out_x0) for (int i = 0; i < NX; ++i) out_x0[(size_t)b * NX + i] = (double)x[i];
    if (out_u0) for (int i = 0; i < NU; ++i) out_u0[(size_t)b * NU + i] = (double)u[i];
    if (x_next) {
      for (int i = 0; i < NX; ++i) s_xn[i] = (T)x_next[(size_t)b * NX + i];
    } else {
      T qdd[NJ], Minv[NJ * NJ], v[NJ][6], a[NJ][6], f[NJ][6], xn[NX];
      if constexpr (MS) integrator_multi_value(d.integrator, x, u, d.gravity, d.dt, xn);
      else {
        forward_dynamics<T, false>(x, x + NJ, u, d.gravity, qdd, Minv, v, a, f);
        integrate(d.integrator, x, qdd, d.dt, xn);
      }
      for (int i = 0; i < NX; ++i) s_xn[i] = xn[i];
    }
    if (out_xnext) for (int i = 0; i < NX; ++i) out_xnext[(size_t)b * NX + i] = (double)s_xn[i];
  }
  __syncthreads();
  for (int base = 0; base < N; base += blockDim.x) {          // chunks left to right: a chunk only reads knots >= its own
    const int k = base + threadIdx.x;
    T xv[NX], uv[NU], mv[3][2 * NM];
    const bool live = k < N;
    const int src = (k + 1 < N) ? k + 1 : N - 1;
    if (live) {
      for (int i = 0; i < NX; ++i) xv[i] = d.x[(size_t)i * K + t0 + src];
      const int usrc = (k + 1 < N - 1) ? k + 1 : N - 2;
      for (int i = 0; i < NU; ++i) uv[i] = d.u[(size_t)i * K + t0 + usrc];
      if (d.lim.any)
        for (int e = 0; e < 2 * NM; ++e) {
          mv[0][e] = d.mu[(size_t)e * K + t0 + src]; mv[1][e] = d.lam[(size_t)e * K + t0 + src]; mv[2][e] = d.phi[(size_t)e * K + t0 + src];
        }
    }
    __syncthreads();
    if (live) {
      for (int i = 0; i < NX; ++i) d.x[(size_t)i * K + t0 + k] = (k == 0) ? s_xn[i] : xv[i];
      if (k < N - 1)
        for (int i = 0; i < NU; ++i) d.u[(size_t)i * K + t0 + k] = uv[i];
      if (d.lim.any)
        for (int e = 0; e < 2 * NM; ++e) {
          const int ci = e % NM;
          const int ty = ci < NJ ? 0 : (ci < NX ? 1 : 2);
          const bool last = (k == N - 1);
          d.mu[(size_t)e * K + t0 + k] = last ? d.mu_init[ty] : mv[0][e];
          d.lam[(size_t)e * K + t0 + k] = last ? T(0) : mv[1][e];
          d.phi[(size_t)e * K + t0 + k] = last ? d.phi_init[ty] : mv[2][e];
        }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0)
    for (int i = 0; i < NX; ++i) d.xs[(size_t)i * d.B + b] = s_xn[i];
}

template <typename T>
__global__ void k_pack_status(Dev<T> d, int* st, double* sc) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= d.B) return;
  int* o = st + (size_t)b * 8;
  o[0] = d.exit_sqp[b]; o[1] = d.exit_soft[b]; o[2] = d.outer_iter[b]; o[3] = d.sqp_iter[b];
  o[4] = d.tot_qp[b]; o[5] = d.tot_pcg[b]; o[6] = d.tot_trials[b]; o[7] = d.trace_rows[b];
  double* s = sc + (size_t)b * 4;
  s[0] = (double)d.J[b]; s[1] = (double)d.c[b]; s[2] = (double)d.merit[b]; s[3] = (double)d.rho[b];
}
template <typename T>
__global__ void k_fetch_trace(Dev<T> d, double* out, int cap) {
  const size_t gt = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t n = (size_t)d.B * cap * TRACE_FIELDS;
  if (gt >= n) return;
  const int f = (int)(gt % TRACE_FIELDS);
  const size_t br = gt / TRACE_FIELDS;
  const int row = (int)(br % cap);
  const int b = (int)(br / cap);
  out[gt] = (row < d.trace_cap) ? (double)d.trace[((size_t)b * d.trace_cap + row) * TRACE_FIELDS + f] : 0.0;
}
}  // namespace b2t
