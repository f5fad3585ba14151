"""Plant oracle: forward dynamics, its gradient and the discrete-time integrator (A, B, x+).

Restates /root/reference/TrajoptPlant.py.  Broadcasts over leading dims of x (..,2n), u (..,n).
"""
import numpy as np

from . import rbd


def forward_dynamics(model, x, u, gravity=-9.81):
    """qdd = Minv(q) (u - rnea(q, qd, 0)).  URDFPlant.forward_dynamics (TrajoptPlant.py:283-299)."""
    n = model.n
    q, qd = x[..., :n], x[..., n:]
    c = rbd.rnea(model, q, qd, None, gravity)[0]
    Minv = rbd.minv(model, q)
    return np.matmul(Minv, (u - c)[..., None])[..., 0]


def forward_dynamics_gradient(model, x, u, gravity=-9.81, return_parts=False):
    """dqdd/d(q,qd,u) = [-Minv dc/dq, -Minv dc/dqd, Minv] with dc evaluated at qdd.
    URDFPlant.forward_dynamics_gradient (TrajoptPlant.py:301-323).  -> (..,n,3n)"""
    n = model.n
    q, qd = x[..., :n], x[..., n:]
    c = rbd.rnea(model, q, qd, None, gravity)[0]
    Minv = rbd.minv(model, q)
    qdd = np.matmul(Minv, (u - c)[..., None])[..., 0]
    dc_du = rbd.rnea_grad(model, q, qd, qdd, gravity)
    df_du = np.matmul(-Minv, dc_du)
    dqdd = np.concatenate([df_du, Minv], axis=-1)
    if return_parts:
        return dqdd, dict(c=c, Minv=Minv, qdd=qdd, dc_du=dc_du)
    return dqdd


def integrator(model, x, u, dt, integrator_type=0, return_gradient=False, gravity=-9.81):
    """TrajoptPlant.integrator (TrajoptPlant.py:83-138), types 0 (Euler) and 1 (semi-implicit Euler).

    Types 2-4 of the reference return Jacobians inconsistent with their own step (2, 3) or raise (4)
    (SURVEY.md 0.9); they are out of scope."""
    n = model.n
    nx = 2 * n
    if integrator_type not in (0, 1):
        raise ValueError("only integrator types 0 (euler) and 1 (semi-implicit euler) are supported")
    if not return_gradient:
        qdd = forward_dynamics(model, x, u, gravity)
        if integrator_type == 0:
            xdot = np.concatenate([x[..., n:], qdd], axis=-1)        # qdd_to_xdot (:61-70)
            return x + dt * xdot
        vkp1 = x[..., n:] + dt * qdd
        qkp1 = x[..., :n] + dt * vkp1
        return np.concatenate([qkp1, vkp1], axis=-1)
    dqdd = forward_dynamics_gradient(model, x, u, gravity)
    bshape = dqdd.shape[:-2]
    top = np.zeros(bshape + (n, 3 * n))
    top[..., :, n:2 * n] = np.eye(n)                                # dqdd_to_dxdot (:72-81)
    if integrator_type == 0:
        dxdot = np.concatenate([top, dqdd], axis=-2)
        A = np.eye(nx) + dt * dxdot[..., :, :nx]
        B = dt * dxdot[..., :, nx:]
        return A, B
    Iz = np.zeros((nx, 3 * n))
    Iz[:, :nx] = np.eye(nx)
    AB = Iz + dt * np.concatenate([top + dt * dqdd, dqdd], axis=-2)
    return AB[..., :, :nx], AB[..., :, nx:]
