// TEST INFRASTRUCTURE ONLY: compiles the per-knot device math of csrc/b2t_core.cuh for the host (g++, no CUDA) so that
// the `-m "not gpu"` tests can check it against the oracle without a GPU.  Never loaded by the product package.
#include "b2t_core.cuh"
using namespace b2t;
extern "C" {
int he_dims(int* nj) { *nj = NJ; return 0; }
// forward dynamics + gradient + integrator for `count` knots; x [count][NX], u [count][NU]
// out: qdd [count][NJ], Minv [count][NJ*NJ], dqdd [count][NJ*3NJ], xn [count][NX], AB [count][NX*NM]
int he_dynamics(int count, const double* x, const double* u, double dt, double gravity, int integrator, double* qdd, double* Minv,
                double* dqdd, double* xn, double* AB) {
  for (int t = 0; t < count; ++t) {
    const double* xt = x + t * NX;
    const double* ut = u + t * NU;
    double v[NJ][6], a[NJ][6], f[NJ][6];
    forward_dynamics<double, true>(xt, xt + NJ, ut, gravity, qdd + t * NJ, Minv + t * NJ * NJ, v, a, f);
    integrate(integrator, xt, qdd + t * NJ, dt, xn + t * NX);
    double* dq = dqdd + t * NDYN;
    for (int colid = 0; colid < 2 * NJ; ++colid) {
      double out[NJ];
      fd_grad_column(xt, xt + NJ, v, a, f, Minv + t * NJ * NJ, gravity, colid % NJ, colid >= NJ, out);
      for (int i = 0; i < NJ; ++i) dq[i * 3 * NJ + colid] = out[i];
    }
    for (int i = 0; i < NJ; ++i)
      for (int j = 0; j < NJ; ++j) dq[i * 3 * NJ + 2 * NJ + j] = Minv[t * NJ * NJ + i * NJ + j];
    build_AB(integrator, dq, dt, AB + t * NX * NM);
  }
  return 0;
}
// integrator types 2 / 3 (the reference's midpoint / rk3): step and [A B] for `count` knots
int he_integrator_multi(int count, int integrator, const double* x, const double* u, double dt, double gravity, double* xn, double* AB) {
  for (int t = 0; t < count; ++t) {
    integrator_multi_value<double>(integrator, x + t * NX, u + t * NU, gravity, dt, xn + t * NX);
    integrator_multi_AB<double>(integrator, x + t * NX, u + t * NU, gravity, dt, AB + t * NX * NM);
  }
  return 0;
}
// qdd by the single-right-hand-side articulated-body solve (forward_dynamics_qdd) for `count` knots
int he_qdd_solve(int count, const double* x, const double* u, double gravity, double* qdd) {
  for (int t = 0; t < count; ++t) forward_dynamics_qdd<double>(x + t * NX, x + t * NX + NJ, u + t * NU, gravity, qdd + t * NJ);
  return 0;
}
// cost value / gradient / hessian and soft terms of `count` knots
int he_cost_mode(int count, int kind, int qf_start, int hess_mode, const double* Q, const double* QF, const double* R, const double* xg,
                 const double* x, const double* u, const int* kidx, const int* terminal, double* val, double* grad, double* hess) {
  CostParams<double> cp{kind, qf_start, hess_mode, Q, QF, R};
  for (int t = 0; t < count; ++t) {
    val[t] = cost_value(cp, x + t * NX, u + t * NU, xg, kidx[t], terminal[t] != 0);
    cost_grad_hess<double, true>(cp, x + t * NX, u + t * NU, xg, kidx[t], terminal[t] != 0, grad + t * NM, hess + t * NM * NM);
  }
  return 0;
}
int he_cost(int count, int kind, int qf_start, const double* Q, const double* QF, const double* R, const double* xg, const double* x,
            const double* u, const int* kidx, const int* terminal, double* val, double* grad, double* hess) {
  CostParams<double> cp{kind, qf_start, 0, Q, QF, R};
  for (int t = 0; t < count; ++t) {
    val[t] = cost_value(cp, x + t * NX, u + t * NU, xg, kidx[t], terminal[t] != 0);
    cost_grad_hess<double, true>(cp, x + t * NX, u + t * NU, xg, kidx[t], terminal[t] != 0, grad + t * NM, hess + t * NM * NM);
  }
  return 0;
}
int he_soft(int count, const int* mode, const double* lb, const double* ub, const double* z, const double* mu, const double* lam,
            const int* terminal, double* val, double* gck) {
  LimitParams<double> lp{1, mode, lb, ub};
  for (int t = 0; t < count; ++t) {
    val[t] = soft_value(lp, z + t * NM, mu + t * 2 * NM, lam + t * 2 * NM, 1, terminal[t] != 0);
    soft_grad(lp, z + t * NM, mu + t * 2 * NM, lam + t * 2 * NM, 1, terminal[t] != 0, gck + t * NM);
  }
  return 0;
}
int he_spd_inverse(int n, double* A) { spd_inverse_inplace(A, n, n); return 0; }
}
extern "C" int he_spd_inverse_packed12(double* a) { spd_inverse_packed<12>(a); return 0; }
extern "C" int he_spd_inverse_packed4(double* a) { spd_inverse_packed<4>(a); return 0; }
