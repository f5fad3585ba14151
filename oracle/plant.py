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


def _xdot(x, qdd, n):
    """qdd_to_xdot (TrajoptPlant.py:61-70): [velocity part of the state it is HANDED ; qdd]."""
    return np.concatenate([x[..., n:], qdd], axis=-1)


def _dxdot(model, x, u, gravity):
    """dqdd_to_dxdot(forward_dynamics_gradient(x, u)) (TrajoptPlant.py:72-81): [[0 I 0]; dqdd]  -> (.., 2n, 3n)"""
    n = model.n
    dqdd = forward_dynamics_gradient(model, x, u, gravity)
    top = np.zeros(dqdd.shape[:-2] + (n, 3 * n))
    top[..., :, n:2 * n] = np.eye(n)
    return np.concatenate([top, dqdd], axis=-2)


def integrator(model, x, u, dt, integrator_type=0, return_gradient=False, gravity=-9.81):
    """TrajoptPlant.integrator (TrajoptPlant.py:83-205), types 0 (Euler), 1 (semi-implicit Euler), 2 (midpoint), 3 (rk3).

    Types 2 and 3 are restated LITERALLY, not as the textbook schemes: every stage derivative is
    qdd_to_xdot(xk, forward_dynamics(point_s, uk)) -- the velocity rows come from xk, not from the stage point (:141-143, :171-175) --
    and rk3 builds B2, B3 from the first stage's gradient dxdot1 (:195, :199), so (A, B) is not the Jacobian of the step.  A drop-in
    has to reproduce exactly that.  Type 4 raises TypeError in the reference's numpy gradient branch (:259: an extra positional xk)."""
    n = model.n
    nx = 2 * n
    if integrator_type == 4:
        raise TypeError("integrator type 4 (rk4): the reference's gradient branch raises (TrajoptPlant.py:259)")
    if integrator_type not in (0, 1, 2, 3):
        raise ValueError("integrator types 0 (euler), 1 (semi-implicit euler), 2 (midpoint), 3 (rk3)")
    if integrator_type == 2:                                                            # :140-168
        xdot1 = _xdot(x, forward_dynamics(model, x, u, gravity), n)
        midpoint = x + 0.5 * dt * xdot1
        if not return_gradient:
            xdot2 = _xdot(x, forward_dynamics(model, midpoint, u, gravity), n)
            return x + dt * xdot2
        dxdot1 = _dxdot(model, x, u, gravity)
        A1 = np.eye(nx) + 0.5 * dt * dxdot1[..., :, :nx]
        B1 = 0.5 * dt * dxdot1[..., :, nx:]
        dxdot2 = _dxdot(model, midpoint, u, gravity)
        A2 = np.eye(nx) + 0.5 * dt * dxdot2[..., :, :nx]
        B2 = 0.5 * dt * dxdot2[..., :, nx:]
        return np.matmul(A2, A1), np.matmul(A2, B1) + B2
    if integrator_type == 3:                                                            # :170-205
        xdot1 = _xdot(x, forward_dynamics(model, x, u, gravity), n)
        point1 = x + 0.5 * dt * xdot1
        xdot2 = _xdot(x, forward_dynamics(model, point1, u, gravity), n)
        point2 = x + 0.75 * dt * xdot2
        if not return_gradient:
            xdot3 = _xdot(x, forward_dynamics(model, point2, u, gravity), n)
            return x + (dt / 9) * (2 * xdot1 + 3 * xdot2 + 4 * xdot3)
        dxdot1 = _dxdot(model, x, u, gravity)
        A1 = np.eye(nx) + 2 / 9 * dt * dxdot1[..., :, :nx]
        B1 = 2 / 9 * dt * dxdot1[..., :, nx:]
        dxdot2 = _dxdot(model, point1, u, gravity)
        A2 = np.eye(nx) + 1 / 3 * dt * dxdot2[..., :, :nx]
        B2 = 1 / 3 * dt * dxdot1[..., :, nx:]
        dxdot3 = _dxdot(model, point2, u, gravity)
        A3 = np.eye(nx) + 4 / 9 * dt * dxdot3[..., :, :nx]
        B3 = 4 / 9 * dt * dxdot1[..., :, nx:]
        A = np.matmul(A3, np.matmul(A2, A1))
        B = np.matmul(A3, np.matmul(A2, B1)) + np.matmul(A3, B2) + B3
        return A, B
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
