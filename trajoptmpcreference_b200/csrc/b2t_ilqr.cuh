// b2t_ilqr.cuh -- iLQR (Gauss-Newton DDP with soft box limits) kernels.  The reference ships no iLQR code (README.md:15-17 and
// the enum MPCSolverMethods.iLQR, TrajoptMPCReference.py:21-27, only name it); the specification is oracle/ilqr.py in this
// repository.  Shares the dynamics, cost, penalty, outer-loop and work-list kernels with the SQP path.
//   k_ilqr_cost      thread / knot      l_x, l_u (g) and the Gauss-Newton Hessian incl. penalty terms (Gh, dense m x m)
//   k_ilqr_backward  thread / instance  Riccati recursion over the knots: gains kff, K, expected-reduction terms dV1, dV2
//   k_ilqr_search    thread / instance  backtracking over alpha: closed-loop rollout through the integrator, cost, accept / fail,
//                                       regularisation schedule and exit tests (same formulas as the SQP path)
//   k_ilqr_rollout0  thread / instance  initial state trajectory = rollout of the given controls
#pragma once
#include "b2t_kernels.cuh"

namespace b2t {

template <typename T>
__global__ void __launch_bounds__(64) k_ilqr_cost(Dev<T> d, const int* list, const int* count) {
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
  for (int i = 0; i < NM; ++i) d.g[(size_t)i * K + t] = g[i];
  for (int i = 0; i < NM * NM; ++i) d.Gh[(size_t)i * K + t] = G[i];
}

// gains are stored in the (otherwise unused) Schur arrays: K_k (NU x NX) in Sd[0 .. NU*NX), kff_k in gam[0 .. NU)
template <typename T>
__global__ void __launch_bounds__(32) k_ilqr_backward(Dev<T> d, const int* list, const int* count) {
  const int slot = blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= *count) return;
  const int b = list[slot];
  const int N = d.N;
  const size_t K = d.K;
  const size_t t0 = (size_t)b * N;
  const T rho = d.rho[b];
  T Vx[NX], Vxx[NX * NX];
  {
    const size_t t = t0 + N - 1;
    for (int i = 0; i < NX; ++i) {
      Vx[i] = d.g[(size_t)i * K + t];
      for (int c = 0; c < NX; ++c) Vxx[i * NX + c] = d.Gh[(size_t)(i * NM + c) * K + t];
    }
  }
  T dV1 = T(0), dV2 = T(0);
  int ok = 1;
  for (int k = N - 2; k >= 0 && ok; --k) {
    const size_t t = t0 + k;
    T AB[NX * NM];
    load_AB(d, t, AB);                                    // [A B], NX x NM
    // VAB = Vxx [A B]  (NX x NM)
    T VAB[NX * NM];
    for (int i = 0; i < NX; ++i)
      for (int c = 0; c < NM; ++c) {
        T acc = T(0);
        for (int r = 0; r < NX; ++r) acc += Vxx[i * NX + r] * AB[r * NM + c];
        VAB[i * NM + c] = acc;
      }
    // Q = H + [A B]^T Vxx [A B]  (NM x NM: blocks Qxx, Qxu; Qux, Quu),  q = g + [A B]^T Vx
    T Q[NM * NM], q[NM];
    for (int a = 0; a < NM; ++a) {
      T accq = d.g[(size_t)a * K + t];
      for (int r = 0; r < NX; ++r) accq += AB[r * NM + a] * Vx[r];
      q[a] = accq;
      for (int c = 0; c < NM; ++c) {
        T acc = d.Gh[(size_t)(a * NM + c) * K + t];
        for (int r = 0; r < NX; ++r) acc += AB[r * NM + a] * VAB[r * NM + c];
        Q[a * NM + c] = acc;
      }
    }
    // Quu + rho I, its inverse (SPD; a non-positive pivot flags failure -> the caller raises rho)
    T Quu[NU * NU], Qi[NU * NU];
    for (int i = 0; i < NU; ++i)
      for (int c = 0; c < NU; ++c) {
        Quu[i * NU + c] = Q[(NX + i) * NM + NX + c] + ((i == c) ? rho : T(0));
        Qi[i * NU + c] = Quu[i * NU + c];
      }
    {   // Cholesky test for positive definiteness (mirrors np.linalg.cholesky in the oracle)
      T L[NU * NU];
      for (int i = 0; i < NU * NU; ++i) L[i] = Quu[i];
      for (int j = 0; j < NU && ok; ++j) {
        T s = L[j * NU + j];
        for (int r = 0; r < j; ++r) s -= L[j * NU + r] * L[j * NU + r];
        if (!(s > T(0))) { ok = 0; break; }
        const T ljj = sqrt(s);
        L[j * NU + j] = ljj;
        for (int i = j + 1; i < NU; ++i) {
          T s2 = L[i * NU + j];
          for (int r = 0; r < j; ++r) s2 -= L[i * NU + r] * L[j * NU + r];
          L[i * NU + j] = s2 / ljj;
        }
      }
    }
    if (!ok) break;
    spd_inverse_inplace(Qi, NU, NU);
    T kff[NU], Kg[NU * NX];
    for (int i = 0; i < NU; ++i) {
      T acc = T(0);
      for (int c = 0; c < NU; ++c) acc += Qi[i * NU + c] * q[NX + c];
      kff[i] = -acc;
      for (int x = 0; x < NX; ++x) {
        T a2 = T(0);
        for (int c = 0; c < NU; ++c) a2 += Qi[i * NU + c] * Q[(NX + c) * NM + x];
        Kg[i * NX + x] = -a2;
      }
    }
    T Qk[NU];        // Quu kff
    for (int i = 0; i < NU; ++i) {
      T acc = T(0);
      for (int c = 0; c < NU; ++c) acc += Quu[i * NU + c] * kff[c];
      Qk[i] = acc;
    }
    T s1 = T(0), s2 = T(0);
    for (int i = 0; i < NU; ++i) { s1 += kff[i] * q[NX + i]; s2 += kff[i] * Qk[i]; }
    dV1 += s1;
    dV2 += T(0.5) * s2;
    // Vx = Qx + K^T Quu kff + K^T Qu + Qux^T kff ;  Vxx = Qxx + K^T Quu K + K^T Qux + Qux^T K
    T QK[NU * NX];   // Quu K
    for (int i = 0; i < NU; ++i)
      for (int x = 0; x < NX; ++x) {
        T acc = T(0);
        for (int c = 0; c < NU; ++c) acc += Quu[i * NU + c] * Kg[c * NX + x];
        QK[i * NX + x] = acc;
      }
    T Vn[NX * NX];
    for (int x = 0; x < NX; ++x) {
      T acc = q[x];
      for (int i = 0; i < NU; ++i) acc += Kg[i * NX + x] * Qk[i] + Kg[i * NX + x] * q[NX + i] + Q[(NX + i) * NM + x] * kff[i];
      Vx[x] = acc;
      for (int y = 0; y < NX; ++y) {
        T a2 = Q[x * NM + y];
        for (int i = 0; i < NU; ++i) a2 += Kg[i * NX + x] * QK[i * NX + y] + Kg[i * NX + x] * Q[(NX + i) * NM + y] + Q[(NX + i) * NM + x] * Kg[i * NX + y];
        Vn[x * NX + y] = a2;
      }
    }
    for (int x = 0; x < NX; ++x)
      for (int y = 0; y < NX; ++y) Vxx[x * NX + y] = T(0.5) * (Vn[x * NX + y] + Vn[y * NX + x]);
    for (int i = 0; i < NU; ++i) {
      d.gam[(size_t)i * K + t] = kff[i];
      for (int x = 0; x < NX; ++x) d.Sd[(size_t)(i * NX + x) * K + t] = Kg[i * NX + x];
    }
  }
  d.D[b] = dV1;
  d.ratio[b] = dV2;
  d.err[b] = ok ? 0 : 1;
  d.tot_qp[b] += 1;
}

// Block-parallel Riccati recursion: one block (ILQR_BW_THREADS threads) per instance, the knots are processed sequentially, every
// matrix product of a knot is spread over the threads (entries of the result), operands staged in shared memory.
// Same arithmetic as k_ilqr_backward (which is kept as the single-thread reference variant, B2T_ILQR_BW=1).
enum { ILQR_BW_THREADS = 128 };
template <typename T>
__global__ void __launch_bounds__(ILQR_BW_THREADS) k_ilqr_backward2(Dev<T> d, const int* list, const int* count) {
  if ((int)blockIdx.x >= *count) return;
  const int b = list[blockIdx.x];
  const int N = d.N;
  const size_t K = d.K;
  const size_t t0 = (size_t)b * N;
  const int tid = threadIdx.x, nt = blockDim.x;
  __shared__ T sAB[NX * NM], sVAB[NX * NM], sQ[NM * NM], sq[NM], sVxx[NX * NX], sVx[NX], sVn[NX * NX];
  __shared__ T sQuu[NU * NU], sAug[NU * 2 * NU], sQi[NU * NU], skff[NU], sKg[NU * NX], sQk[NU], sQK[NU * NX], sdq[NDYN];
  __shared__ int s_ok;
  const T rho = d.rho[b];
  {
    const size_t t = t0 + N - 1;
    for (int i = tid; i < NX; i += nt) sVx[i] = d.g[(size_t)i * K + t];
    for (int e = tid; e < NX * NX; e += nt) sVxx[e] = d.Gh[(size_t)((e / NX) * NM + (e % NX)) * K + t];
    if (tid == 0) s_ok = 1;
  }
  T dV1 = T(0), dV2 = T(0);
  __syncthreads();
  for (int k = N - 2; k >= 0; --k) {
    const size_t t = t0 + k;
    if (d.integrator < 2)
      for (int i = tid; i < NDYN; i += nt) sdq[i] = d.dyn[(size_t)i * K + t];
    __syncthreads();
    // [A B] (build_AB, entry-parallel)
    for (int e = tid; e < NX * NM; e += nt) {
      const int r = e / NM, c = e % NM;
      T v;
      if (d.integrator >= 2) v = d.ABf[(size_t)e * K + t];      // midpoint / rk3: the full [A B] of k_ab_multi
      else if (d.integrator == 0) {
        if (r < NJ) v = ((c == r) ? T(1) : T(0)) + ((c == NJ + r) ? d.dt : T(0));
        else v = d.dt * sdq[(r - NJ) * 3 * NJ + c] + ((c == r) ? T(1) : T(0));
      } else {
        const int a = (r < NJ) ? r : r - NJ;
        const T dd = sdq[a * 3 * NJ + c];
        if (r < NJ) v = d.dt * (((c == NJ + a) ? T(1) : T(0)) + d.dt * dd) + ((c == r) ? T(1) : T(0));
        else v = d.dt * dd + ((c == r) ? T(1) : T(0));
      }
      sAB[e] = v;
    }
    __syncthreads();
    for (int e = tid; e < NX * NM; e += nt) {            // VAB = Vxx [A B]
      const int i = e / NM, c = e % NM;
      T acc = T(0);
      for (int r = 0; r < NX; ++r) acc += sVxx[i * NX + r] * sAB[r * NM + c];
      sVAB[e] = acc;
    }
    __syncthreads();
    for (int e = tid; e < NM * NM + NM; e += nt) {       // Q = H + [A B]^T VAB ; q = g + [A B]^T Vx
      if (e < NM * NM) {
        const int a = e / NM, c = e % NM;
        T acc = d.Gh[(size_t)e * K + t];
        for (int r = 0; r < NX; ++r) acc += sAB[r * NM + a] * sVAB[r * NM + c];
        sQ[e] = acc;
      } else {
        const int a = e - NM * NM;
        T acc = d.g[(size_t)a * K + t];
        for (int r = 0; r < NX; ++r) acc += sAB[r * NM + a] * sVx[r];
        sq[a] = acc;
      }
    }
    __syncthreads();
    for (int e = tid; e < NU * NU; e += nt) {
      const int i = e / NU, c = e % NU;
      const T v = sQ[(NX + i) * NM + NX + c] + ((i == c) ? rho : T(0));
      sQuu[e] = v;
      sAug[i * 2 * NU + c] = v;
      sAug[i * 2 * NU + NU + c] = (i == c) ? T(1) : T(0);
    }
    __syncthreads();
    // Gauss-Jordan on [Quu | I] without pivoting (SPD), row-parallel within warp 0; a non-positive pivot flags failure
    if (tid < 32) {
      for (int p = 0; p < NU; ++p) {
        const T piv = sAug[p * 2 * NU + p];
        if (!(piv > T(0))) { if (tid == 0) s_ok = 0; break; }
        __syncwarp();
        if (tid < 2 * NU) sAug[p * 2 * NU + tid] = sAug[p * 2 * NU + tid] / piv;
        __syncwarp();
        if (tid < NU && tid != p) {
          const T fct = sAug[tid * 2 * NU + p];
          for (int c = 0; c < 2 * NU; ++c) sAug[tid * 2 * NU + c] -= fct * sAug[p * 2 * NU + c];
        }
        __syncwarp();
      }
    }
    __syncthreads();
    if (!s_ok) break;
    for (int e = tid; e < NU * NU; e += nt) sQi[e] = sAug[(e / NU) * 2 * NU + NU + (e % NU)];
    __syncthreads();
    for (int e = tid; e < NU + NU * NX; e += nt) {       // kff = -Qi Qu ; K = -Qi Qux
      if (e < NU) {
        T acc = T(0);
        for (int c = 0; c < NU; ++c) acc += sQi[e * NU + c] * sq[NX + c];
        skff[e] = -acc;
      } else {
        const int i = (e - NU) / NX, x = (e - NU) % NX;
        T acc = T(0);
        for (int c = 0; c < NU; ++c) acc += sQi[i * NU + c] * sQ[(NX + c) * NM + x];
        sKg[i * NX + x] = -acc;
      }
    }
    __syncthreads();
    for (int e = tid; e < NU + NU * NX; e += nt) {       // Qk = Quu kff ; QK = Quu K
      if (e < NU) {
        T acc = T(0);
        for (int c = 0; c < NU; ++c) acc += sQuu[e * NU + c] * skff[c];
        sQk[e] = acc;
      } else {
        const int i = (e - NU) / NX, x = (e - NU) % NX;
        T acc = T(0);
        for (int c = 0; c < NU; ++c) acc += sQuu[i * NU + c] * sKg[c * NX + x];
        sQK[i * NX + x] = acc;
      }
    }
    __syncthreads();
    if (tid == 0) {
      T s1 = T(0), s2 = T(0);
      for (int i = 0; i < NU; ++i) { s1 += skff[i] * sq[NX + i]; s2 += skff[i] * sQk[i]; }
      dV1 += s1;
      dV2 += T(0.5) * s2;
    }
    for (int e = tid; e < NX * NX + NX; e += nt) {       // Vx, Vxx (unsymmetrised)
      if (e < NX * NX) {
        const int x = e / NX, y = e % NX;
        T a2 = sQ[x * NM + y];
        for (int i = 0; i < NU; ++i) a2 += sKg[i * NX + x] * sQK[i * NX + y] + sKg[i * NX + x] * sQ[(NX + i) * NM + y] + sQ[(NX + i) * NM + x] * sKg[i * NX + y];
        sVn[e] = a2;
      } else {
        const int x = e - NX * NX;
        T acc = sq[x];
        for (int i = 0; i < NU; ++i) acc += sKg[i * NX + x] * sQk[i] + sKg[i * NX + x] * sq[NX + i] + sQ[(NX + i) * NM + x] * skff[i];
        sVx[x] = acc;
      }
    }
    for (int e = tid; e < NU + NU * NX; e += nt) {       // gains to global (kff -> gam, K -> Sd)
      if (e < NU) d.gam[(size_t)e * K + t] = skff[e];
      else d.Sd[(size_t)(e - NU) * K + t] = sKg[e - NU];
    }
    __syncthreads();
    for (int e = tid; e < NX * NX; e += nt) { const int x = e / NX, y = e % NX; sVxx[e] = T(0.5) * (sVn[x * NX + y] + sVn[y * NX + x]); }
    __syncthreads();
  }
  if (tid == 0) {
    d.D[b] = dV1;
    d.ratio[b] = dV2;
    d.err[b] = s_ok ? 0 : 1;
    d.tot_qp[b] += 1;
  }
}

template <typename T>
__global__ void __launch_bounds__(32) k_ilqr_rollout0(Dev<T> d) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= d.B) return;
  const int N = d.N;
  const size_t K = d.K;
  const size_t t0 = (size_t)b * N;
  T x[NX];
  for (int i = 0; i < NX; ++i) x[i] = d.x[(size_t)i * K + t0];
  for (int k = 0; k < N - 1; ++k) {
    T u[NU], qdd[NJ], Minv[NJ * NJ], v[NJ][6], a[NJ][6], f[NJ][6], xn[NX];
    for (int i = 0; i < NU; ++i) u[i] = d.u[(size_t)i * K + t0 + k];
    if (d.integrator >= 2) integrator_multi_value(d.integrator, x, u, d.gravity, d.dt, xn);
    else {
      forward_dynamics<T, false>(x, x + NJ, u, d.gravity, qdd, Minv, v, a, f);
      integrate(d.integrator, x, qdd, d.dt, xn);
    }
    for (int i = 0; i < NX; ++i) { x[i] = xn[i]; d.x[(size_t)i * K + t0 + k + 1] = xn[i]; }
  }
}

enum { ILQR_MAX_KNOTS = 512 };

template <typename T>
__global__ void __launch_bounds__(32) k_ilqr_search(Dev<T> d, Opts<T> o) {
  const int slot = blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= *d.n_act) return;
  const int b = d.act[slot];
  const int N = d.N;
  const size_t K = d.K;
  const size_t t0 = (size_t)b * N;
  T xg[NX];
  for (int i = 0; i < NX; ++i) xg[i] = d.xg[(size_t)i * d.B + b];
  const T dV1 = d.D[b], dV2 = d.ratio[b];
  bool error = d.err[b] != 0;
  T alpha = T(1);
  int ls = 0;
  T delta_J = T(0);
  T soft[ILQR_MAX_KNOTS];
  while (!error) {
    T z[NM], Jc = T(0);
    for (int i = 0; i < NX; ++i) z[i] = d.x[(size_t)i * K + t0];            // x_0 is fixed
    for (int k = 0; k < N; ++k) {
      const size_t t = t0 + k;
      const bool terminal = (k == N - 1);
      for (int i = 0; i < NX; ++i) d.xn[(size_t)i * K + t] = z[i];
      if (!terminal) {
        for (int i = 0; i < NU; ++i) {
          T acc = d.u[(size_t)i * K + t] + alpha * d.gam[(size_t)i * K + t];
          for (int x = 0; x < NX; ++x) acc += d.Sd[(size_t)(i * NX + x) * K + t] * (z[x] - d.x[(size_t)x * K + t]);
          z[NX + i] = acc;
          d.un[(size_t)i * K + t] = acc;
        }
      } else {
        for (int i = 0; i < NU; ++i) z[NX + i] = T(0);
      }
      Jc += cost_value(d.cost, z, z + NX, xg, k, terminal);
      soft[k] = d.lim.any ? soft_value(d.lim, z, d.mu + t, d.lam + t, K, terminal) : T(0);
      if (!terminal) {
        T qdd[NJ], Minv[NJ * NJ], v[NJ][6], a[NJ][6], f[NJ][6], xn[NX];
        if (d.integrator >= 2) integrator_multi_value(d.integrator, z, z + NX, d.gravity, d.dt, xn);
        else {
          forward_dynamics<T, false>(z, z + NJ, z + NX, d.gravity, qdd, Minv, v, a, f);
          integrate(d.integrator, z, qdd, d.dt, xn);
        }
        for (int i = 0; i < NX; ++i) z[i] = xn[i];
      }
    }
    T Jn = Jc;
    if (d.lim.any)
      for (int k = 0; k < N; ++k) Jn += soft[k];
    delta_J = d.J[b] - Jn;
    const T expected = -(alpha * dV1 + alpha * alpha * dV2);
    const T ratio = delta_J / expected;
    d.tot_trials[b] += 1;
    const bool finite = (Jn == Jn) && (fabs(Jn) < T(1e300));
    if (expected > T(0) && finite && ratio >= o.er_min && ratio <= o.er_max) {
      for (int k = 0; k < N; ++k) {
        for (int i = 0; i < NX; ++i) d.x[(size_t)i * K + t0 + k] = d.xn[(size_t)i * K + t0 + k];
        if (k < N - 1)
          for (int i = 0; i < NU; ++i) d.u[(size_t)i * K + t0 + k] = d.un[(size_t)i * K + t0 + k];
      }
      d.J[b] = Jn;
      d.merit[b] = Jn;
      const T drho = fmin(d.drho[b] / o.rho_factor, T(1) / o.rho_factor);
      d.drho[b] = drho; d.rho[b] = fmax(d.rho[b] * drho, o.rho_min);
      trace_row(d, b, ls, alpha, dV1, ratio, 0, 1);
      break;
    } else if (alpha > o.alpha_min) {
      alpha *= o.alpha_factor;
      ls += 1;
    } else {
      error = true;
      trace_row(d, b, ls, alpha, dV1, ratio, 0, 0);
    }
  }
  // check_for_exit_or_error (TrajoptMPCReference.py:463-481)
  bool exit_flag = false;
  if (error) {
    const T drho = fmax(d.drho[b] * o.rho_factor, o.rho_factor);
    const T rho = fmax(d.rho[b] * drho, o.rho_min);
    d.drho[b] = drho; d.rho[b] = rho;
    if (rho > o.rho_max) { d.exit_sqp[b] = 2; exit_flag = true; }
  } else if (delta_J < o.tol_sqp) {
    d.exit_sqp[b] = 1; exit_flag = true;
  }
  if (d.sqp_iter[b] == o.max_iter_sqp - 1) { d.exit_sqp[b] = 3; exit_flag = true; }
  else d.sqp_iter[b] += 1;
  if (exit_flag) d.phase[b] = PH_OUTER;
  d.deltaJ[b] = delta_J;
  d.ls_iter[b] = ls;
  d.alpha[b] = alpha;
}

// -----------------------------------------------------------------------------------------------------------------
// k_ilqr_search2: the same line search with every step length tried AT ONCE.  ILQR_LPI lanes per instance, lane i rolls the closed
// loop out with alpha_i = alpha_factor^i (computed by the same repeated multiplication as the sequential loop); the accepted trial is
// the lowest accepted index -- exactly the trial the sequential loop of k_ilqr_search stops at -- and the counters advance by that
// index + 1, so results, iteration and trial counts are identical; only the latency drops from (trials) to one rollout.
// Trial trajectories and per-knot penalty values go to `scratch` [(NM + 1)][K][ILQR_LPI] (lanes of an instance are contiguous: one
// 128-byte line per element); the winner's trajectory is copied to x, u by all lanes of the instance.
// -----------------------------------------------------------------------------------------------------------------
enum { ILQR_LPI = 16 };

template <typename T>
__global__ void __launch_bounds__(64) k_ilqr_search2(Dev<T> d, Opts<T> o, T* scratch, int max_trials) {
  const int gl = blockIdx.x * blockDim.x + threadIdx.x;
  const int slot = gl / ILQR_LPI, lane = gl % ILQR_LPI;
  const bool slot_ok = slot < *d.n_act;
  const int b = slot_ok ? d.act[slot] : 0;
  const int N = d.N;
  const size_t K = d.K;
  const size_t t0 = (size_t)b * N;
  const bool error0 = d.err[b] != 0;
  const bool active = slot_ok && !error0 && lane < max_trials;
  T alpha = T(1);
  for (int i = 0; i < lane; ++i) alpha *= o.alpha_factor;
  const T dV1 = d.D[b], dV2 = d.ratio[b];
  auto SC = [&](int e, size_t t) -> T& { return scratch[((size_t)e * K + t) * ILQR_LPI + lane]; };
  T delta_J = T(0), ratio = T(0), Jn = T(0);
  bool accept = false;
  if (active) {
    T xg[NX];
    for (int i = 0; i < NX; ++i) xg[i] = d.xg[(size_t)i * d.B + b];
    T z[NM], Jc = T(0);
    for (int i = 0; i < NX; ++i) z[i] = d.x[(size_t)i * K + t0];            // x_0 is fixed
    for (int k = 0; k < N; ++k) {
      const size_t t = t0 + k;
      const bool terminal = (k == N - 1);
      for (int i = 0; i < NX; ++i) SC(i, t) = z[i];
      if (!terminal) {
        for (int i = 0; i < NU; ++i) {
          T acc = d.u[(size_t)i * K + t] + alpha * d.gam[(size_t)i * K + t];
          for (int x = 0; x < NX; ++x) acc += d.Sd[(size_t)(i * NX + x) * K + t] * (z[x] - d.x[(size_t)x * K + t]);
          z[NX + i] = acc;
          SC(NX + i, t) = acc;
        }
      } else {
        for (int i = 0; i < NU; ++i) z[NX + i] = T(0);
      }
      Jc += cost_value(d.cost, z, z + NX, xg, k, terminal);
      if (d.lim.any) SC(NM, t) = soft_value(d.lim, z, d.mu + t, d.lam + t, K, terminal);
      if (!terminal) {
        T qdd[NJ], Minv[NJ * NJ], v[NJ][6], a[NJ][6], f[NJ][6], xn[NX];
        if (d.integrator >= 2) integrator_multi_value(d.integrator, z, z + NX, d.gravity, d.dt, xn);
        else {
          forward_dynamics<T, false>(z, z + NJ, z + NX, d.gravity, qdd, Minv, v, a, f);
          integrate(d.integrator, z, qdd, d.dt, xn);
        }
        for (int i = 0; i < NX; ++i) z[i] = xn[i];
      }
    }
    Jn = Jc;
    if (d.lim.any)
      for (int k = 0; k < N; ++k) Jn += SC(NM, t0 + k);                      // costs first, then the penalties knot by knot (totalCost :296-310)
    delta_J = d.J[b] - Jn;
    const T expected = -(alpha * dV1 + alpha * alpha * dV2);
    ratio = delta_J / expected;
    const bool finite = (Jn == Jn) && (fabs(Jn) < T(1e300));
    accept = expected > T(0) && finite && ratio >= o.er_min && ratio <= o.er_max;
  }
  const unsigned grp_shift = (threadIdx.x & 31) / ILQR_LPI * ILQR_LPI;
  const unsigned votes = (__ballot_sync(0xffffffffu, accept) >> grp_shift) & ((1u << ILQR_LPI) - 1u);
  __syncwarp();                       // the winner's scratch stores are ordered before the other lanes' loads below
  if (!slot_ok) return;
  const int win = votes ? __ffs((int)votes) - 1 : -1;
  const int last = error0 ? -1 : (win >= 0 ? win : max_trials - 1);           // the trial the sequential loop ends on
  if (win >= 0) {
    for (int idx = lane; idx < N * NM; idx += ILQR_LPI) {
      const int e = idx / N, k = idx % N;
      if (e >= NX && k == N - 1) continue;
      const T val = scratch[((size_t)e * K + t0 + k) * ILQR_LPI + win];
      if (e < NX) d.x[(size_t)e * K + t0 + k] = val; else d.u[(size_t)(e - NX) * K + t0 + k] = val;
    }
  }
  if (lane != (last < 0 ? 0 : last)) return;
  bool error = error0;
  if (!error0) {
    d.tot_trials[b] += last + 1;
    if (win >= 0) {
      d.J[b] = Jn;
      d.merit[b] = Jn;
      const T drho = fmin(d.drho[b] / o.rho_factor, T(1) / o.rho_factor);
      d.drho[b] = drho; d.rho[b] = fmax(d.rho[b] * drho, o.rho_min);
      trace_row(d, b, last, alpha, dV1, ratio, 0, 1);
    } else {
      error = true;
      trace_row(d, b, last, alpha, dV1, ratio, 0, 0);
    }
  }
  // check_for_exit_or_error (TrajoptMPCReference.py:463-481)
  bool exit_flag = false;
  if (error) {
    const T drho = fmax(d.drho[b] * o.rho_factor, o.rho_factor);
    const T rho = fmax(d.rho[b] * drho, o.rho_min);
    d.drho[b] = drho; d.rho[b] = rho;
    if (rho > o.rho_max) { d.exit_sqp[b] = 2; exit_flag = true; }
  } else if (delta_J < o.tol_sqp) {
    d.exit_sqp[b] = 1; exit_flag = true;
  }
  if (d.sqp_iter[b] == o.max_iter_sqp - 1) { d.exit_sqp[b] = 3; exit_flag = true; }
  else d.sqp_iter[b] += 1;
  if (exit_flag) d.phase[b] = PH_OUTER;
  d.deltaJ[b] = delta_J;
  d.ls_iter[b] = last < 0 ? 0 : last;
  d.alpha[b] = last < 0 ? T(1) : alpha;
}

}  // namespace b2t
