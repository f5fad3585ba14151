"""Receding-horizon MPC loop around the SQP oracle -- UNPINNED (the reference names MPC, README.md:15-17 / MPCSolverMethods
TrajoptMPCReference.py:21-27, but has no loop).  Specification of trajoptmpcreference_b200's `mpc_batch` / `b2t_mpc_shift`:
  repeat `steps` times: solve from the warm start; apply u_0; next initial state = plant integrator(x_0, u_0);
  shift x and u one knot to the left (last knot repeated); x_0 <- next state; multipliers shifted one column, last column
  re-initialised (the intent of TrajoptConstraint.shift_soft_constraint_constants, TrajoptConstraint.py:168-176)."""
import numpy as np

from . import plant as _plant
from . import sqp as _sqp


def shift_limits(cons):
    if cons is None:
        return
    for lim in cons.limits.values():
        for arr, init in ((lim.mu, lim.mu_init), (lim.lam, 0.0), (lim.phi, lim.phi_init)):
            arr[:, :-1] = arr[:, 1:]
            arr[:, -1] = init


def mpc(model, cost, cons, x_start, N, dt, steps, method="PCG-SS", options=None, integrator_type=0, gravity=-9.81):
    nx = x_start.shape[0]
    nu = model.n
    x = np.repeat(np.asarray(x_start, dtype=np.float64)[:, None], N, axis=1)
    u = np.zeros((nu, N - 1))
    xc = [x[:, 0].copy()]; ua = []; its = []; Js = []
    for _ in range(steps):
        r = _sqp.sqp(model, cost, cons, x, u, N, dt, method, options, integrator_type, gravity)
        x, u = r["x"], r["u"]
        its.append(r["sqp_iter"]); Js.append(r["J"])
        xn = _plant.integrator(model, x[:, 0], u[:, 0], dt, integrator_type, False, gravity)
        ua.append(u[:, 0].copy()); xc.append(xn.copy())
        x = np.concatenate([x[:, 1:], x[:, -1:]], axis=1); x[:, 0] = xn
        u = np.concatenate([u[:, 1:], u[:, -1:]], axis=1)
        shift_limits(cons)
    return dict(x_closed=np.array(xc).T, u_applied=np.array(ua).T, sqp_iter=its, J=Js)
