"""iLQR oracle (Gauss-Newton DDP with soft box limits) -- UNPINNED: the reference ships no iLQR code.

The reference only names iLQR (README.md:15-17, `MPCSolverMethods.iLQR` TrajoptMPCReference.py:21-27) and shares its
option names with SQP (`*_SQP_DDP`, set_default_options :98-109).  This file is the specification this repository owns
(SURVEY.md appendix C); the CUDA kernels (csrc/b2t_ilqr.cuh) are tested against it.

Algorithm (per instance).  Same plant `integrator` (A_k, B_k), same cost gradient / Gauss-Newton Hessian, same penalty
terms (g_k += gck, G_k += gck gck^T) and the same outer soft-constraint loop as SQP (sqp.py).
  start   : x_0 = x[:,0]; the initial state trajectory is the ROLLOUT of u from x_0 (dynamically feasible).
  backward: V_x = l_x(N-1), V_xx = l_xx(N-1);  for k = N-2 .. 0
              Q_x = l_x + A^T V_x,  Q_u = l_u + B^T V_x,  Q_xx = l_xx + A^T V_xx A,
              Q_ux = l_ux + B^T V_xx A,  Q_uu = l_uu + B^T V_xx B + rho I            (rho: the SQP regularisation schedule)
              kff = -Q_uu^-1 Q_u,  K = -Q_uu^-1 Q_ux
              V_x = Q_x + K^T Q_uu kff + K^T Q_u + Q_ux^T kff,  V_xx = Q_xx + K^T Q_uu K + K^T Q_ux + Q_ux^T K  (symmetrised)
              dV1 += kff^T Q_u,  dV2 += 1/2 kff^T Q_uu kff
  forward : x_0 fixed,  u_k = ubar_k + alpha kff_k + K_k (x_k - xbar_k),  x_{k+1} = integrator(x_k, u_k)
  search  : alpha = 1, halved (alpha_factor) while alpha > alpha_min;  accept iff expected = -(alpha dV1 + alpha^2 dV2) > 0 and
            expected_reduction_min <= (J - J_new) / expected <= expected_reduction_max
  accept  : reduce_regularization (:457-461);   fail: check_for_exit_or_error's rho increase (:463-471)
  exits   : identical codes to SQP: 1 (delta_J < exit_tolerance_SQP_DDP, negative included), 2 (rho > rho_max), 3 (max iterations)
  outer   : check_and_update_soft_constraints (:483-508), J re-evaluated with the new penalties, rho reset.
"""
import numpy as np

from . import plant as _plant
from .sqp import default_options, total_cost


def rollout(model, x0, U, dt, integrator_type=0, gravity=-9.81):
    N = U.shape[0] + 1
    X = np.zeros((N, x0.shape[0]))
    X[0] = x0
    for k in range(N - 1):
        X[k + 1] = _plant.integrator(model, X[k], U[k], dt, integrator_type, False, gravity)
    return X


def backward_pass(A, B, g, H, rho, nx):
    """Returns kff (N-1,nu), K (N-1,nu,nx), dV1, dV2, ok."""
    N = g.shape[0]
    nu = B.shape[2]
    Vx = g[N - 1, :nx].copy()
    Vxx = H[N - 1, :nx, :nx].copy()
    kff = np.zeros((N - 1, nu)); K = np.zeros((N - 1, nu, nx))
    dV1 = 0.0; dV2 = 0.0
    for k in range(N - 2, -1, -1):
        lx, lu = g[k, :nx], g[k, nx:]
        lxx, luu, lux = H[k, :nx, :nx], H[k, nx:, nx:], H[k, nx:, :nx]
        Qx = lx + A[k].T @ Vx
        Qu = lu + B[k].T @ Vx
        VA = Vxx @ A[k]
        VB = Vxx @ B[k]
        Qxx = lxx + A[k].T @ VA
        Qux = lux + B[k].T @ VA
        Quu = luu + B[k].T @ VB + rho * np.eye(nu)
        try:
            np.linalg.cholesky(Quu)
        except np.linalg.LinAlgError:
            return kff, K, dV1, dV2, False
        Qinv = np.linalg.inv(Quu)
        kff[k] = -Qinv @ Qu
        K[k] = -Qinv @ Qux
        dV1 += float(kff[k] @ Qu)
        dV2 += 0.5 * float(kff[k] @ (Quu @ kff[k]))
        Vx = Qx + K[k].T @ (Quu @ kff[k]) + K[k].T @ Qu + Qux.T @ kff[k]
        Vxx = Qxx + K[k].T @ (Quu @ K[k]) + K[k].T @ Qux + Qux.T @ K[k]
        Vxx = 0.5 * (Vxx + Vxx.T)
    return kff, K, dV1, dV2, True


def ilqr(model, cost, cons, x, u, N, dt, options=None, integrator_type=0, gravity=-9.81):
    o = default_options(options)
    U = np.array(u, dtype=np.float64).T.copy()
    x0 = np.array(x, dtype=np.float64)[:, 0].copy()
    nx = x0.shape[0]
    X = rollout(model, x0, U, dt, integrator_type, gravity)
    exit_sqp = exit_soft = 0
    outer = 0
    trace = []
    total_iters = 0
    while True:
        rho = o["rho_init_SQP_DDP"]; drho = 1
        J = total_cost(cost, cons, X, U)
        it = 0
        while True:
            A, B = _plant.integrator(model, X[:N - 1], U, dt, integrator_type, True, gravity)
            g = cost.gradients(X, U)
            H = cost.hessians(X, U)
            if cons is not None and cons.any():
                gck = cons.gradients(X, U)
                g = g + gck
                H = H + gck[:, :, None] * gck[:, None, :]
            kff, K, dV1, dV2, ok = backward_pass(A, B, g, H, rho, nx)
            total_iters += 1
            alpha = 1.0
            error = not ok
            ls = 0
            delta_J = 0.0
            while ok:
                Xn = np.zeros_like(X); Un = np.zeros_like(U)
                Xn[0] = x0
                for k in range(N - 1):
                    Un[k] = U[k] + alpha * kff[k] + K[k] @ (Xn[k] - X[k])
                    Xn[k + 1] = _plant.integrator(model, Xn[k], Un[k], dt, integrator_type, False, gravity)
                J_new = total_cost(cost, cons, Xn, Un)
                delta_J = J - J_new
                expected = -(alpha * dV1 + alpha * alpha * dV2)
                ratio = delta_J / expected if expected != 0 else np.inf
                if expected > 0 and np.isfinite(J_new) and ratio >= o["expected_reduction_min_SQP_DDP"] and ratio <= o["expected_reduction_max_SQP_DDP"]:
                    X, U, J = Xn, Un, J_new
                    drho = min(drho / o["rho_factor_SQP_DDP"], 1 / o["rho_factor_SQP_DDP"])
                    rho = max(rho * drho, o["rho_min_SQP_DDP"])
                    trace.append(dict(outer_iteration=outer, iteration=it, line_search_iteration=ls, alpha=alpha, rho=rho, J=J,
                                      reduction_ratio=float(ratio), succeeded_line_search=True))
                    break
                elif alpha > o["alpha_min_SQP_DDP"]:
                    alpha *= o["alpha_factor_SQP_DDP"]
                    ls += 1
                else:
                    error = True
                    trace.append(dict(outer_iteration=outer, iteration=it, line_search_iteration=ls, alpha=alpha, rho=rho, J=J,
                                      reduction_ratio=float(ratio), succeeded_line_search=False))
                    break
            exit_flag = False
            if error:
                drho = max(drho * o["rho_factor_SQP_DDP"], o["rho_factor_SQP_DDP"])
                rho = max(rho * drho, o["rho_min_SQP_DDP"])
                if rho > o["rho_max_SQP_DDP"]:
                    exit_sqp = 2; exit_flag = True
            elif delta_J < o["exit_tolerance_SQP_DDP"]:
                exit_sqp = 1; exit_flag = True
            if it == o["max_iter_SQP_DDP"] - 1:
                exit_sqp = 3; exit_flag = True
            else:
                it += 1
            if exit_flag:
                break
        exit_flag = False
        max_c = cons.max_value(X, U) if (cons is not None and cons.any()) else 0
        if max_c < o["exit_tolerance_softConstraints"]:
            exit_soft = 1; exit_flag = True
        if outer == o["max_iter_softConstraints"] - 1:
            exit_soft = 2; exit_flag = True
        else:
            outer += 1
        if not exit_flag:
            if cons.update(X, U):
                exit_soft = 3; exit_flag = True
        if exit_flag:
            break
    return dict(x=X.T.copy(), u=U.T.copy(), exit_sqp=exit_sqp, exit_soft=exit_soft, outer_iter=outer, sqp_iter=it, trace=trace, J=J,
                total_iters=total_iters, total_trials=sum(t["line_search_iteration"] + 1 for t in trace))
