"""Rigid-body dynamics oracle: RNEA, analytical Minv, dRNEA, planar end-effector kinematics.

Restates /root/reference/GRiD/RBDReference/RBDReference.py.  All functions broadcast over leading batch
dimensions: q, qd, qdd, u have shape (..., n).  A "model" is the dict produced by
`trajoptmpcreference_b200.model.extract_model` (or loaded from tests/golden/models.json):
  n, parent[n], S[n][6], jtype[n] ('revolute'|'prismatic'),
  X0,Xa,Xb [n][6][6]   : X_j(t) = X0 + f1(t) Xa + f2(t) Xb, (f1,f2)=(cos t, sin t) revolute, (t, 0) prismatic
  H0,Ha,Hb [n][4][4]   : homogeneous transform, same basis (RBDReference.py:123-148 uses it for the end effector)
  I [n][6][6]          : spatial inertias
"""
import numpy as np


class Model:
    def __init__(self, d):
        self.n = int(d["n"])
        self.parent = [int(p) for p in d["parent"]]
        self.jtype = list(d["jtype"])
        self.S = np.asarray(d["S"], dtype=np.float64).reshape(self.n, 6)
        self.X0 = np.asarray(d["X0"], dtype=np.float64).reshape(self.n, 6, 6)
        self.Xa = np.asarray(d["Xa"], dtype=np.float64).reshape(self.n, 6, 6)
        self.Xb = np.asarray(d["Xb"], dtype=np.float64).reshape(self.n, 6, 6)
        self.H0 = np.asarray(d["H0"], dtype=np.float64).reshape(self.n, 4, 4)
        self.Ha = np.asarray(d["Ha"], dtype=np.float64).reshape(self.n, 4, 4)
        self.Hb = np.asarray(d["Hb"], dtype=np.float64).reshape(self.n, 4, 4)
        self.I = np.asarray(d["I"], dtype=np.float64).reshape(self.n, 6, 6)
        self.subtree = [[j for j in range(self.n) if self._is_desc(j, i)] for i in range(self.n)]

    def _is_desc(self, j, i):
        while j != -1:
            if j == i:
                return True
            j = self.parent[j]
        return False

    def ancestors(self, j):
        out = []
        j = self.parent[j]
        while j != -1:
            out.append(j)
            j = self.parent[j]
        return out

    def leaves(self):
        return [j for j in range(self.n) if len(self.subtree[j]) == 1]

    def basis(self, j, t):
        if self.jtype[j] == "revolute":
            return np.cos(t), np.sin(t)
        return t, np.zeros_like(t)

    def dbasis(self, j, t):
        if self.jtype[j] == "revolute":
            return -np.sin(t), np.cos(t)
        return np.ones_like(t), np.zeros_like(t)

    def X(self, j, t):
        """6x6 motion transform of joint j at position t (..,) -> (..,6,6).  Robot.get_Xmat_Func_by_id (Robot.py:218)."""
        f1, f2 = self.basis(j, np.asarray(t, dtype=np.float64))
        return self.X0[j] + f1[..., None, None] * self.Xa[j] + f2[..., None, None] * self.Xb[j]

    def H(self, j, t):
        """4x4 homogeneous transform (Joint.py:92-97, Robot.py:268)."""
        f1, f2 = self.basis(j, np.asarray(t, dtype=np.float64))
        return self.H0[j] + f1[..., None, None] * self.Ha[j] + f2[..., None, None] * self.Hb[j]

    def dH(self, j, t):
        """d/dt of the homogeneous transform (Joint.py:99, Robot.py:318)."""
        f1, f2 = self.dbasis(j, np.asarray(t, dtype=np.float64))
        return f1[..., None, None] * self.Ha[j] + f2[..., None, None] * self.Hb[j]

    def ddH(self, j, t):
        """second derivative of the homogeneous transform (revolute: -(c Ha + s Hb); prismatic: 0)"""
        t = np.asarray(t, dtype=np.float64)
        if self.jtype[j] == "revolute":
            return -np.cos(t)[..., None, None] * self.Ha[j] - np.sin(t)[..., None, None] * self.Hb[j]
        return np.zeros(t.shape + (4, 4))


    def dddH(self, j, t):
        """third derivative of the homogeneous transform (revolute: -dH; prismatic: 0)"""
        t = np.asarray(t, dtype=np.float64)
        if self.jtype[j] == "revolute":
            return -self.dH(j, t)
        return np.zeros(t.shape + (4, 4))


def _mv(M, v):
    """(..,a,b) @ (..,b) -> (..,a)"""
    return np.matmul(M, v[..., None])[..., 0]


def _mtv(M, v):
    """M^T v"""
    return np.matmul(np.swapaxes(M, -1, -2), v[..., None])[..., 0]


def crm(v):
    """Spatial motion cross-product operator v x (RBDReference.py:13-34), v (..,6) -> (..,6,6)."""
    z = np.zeros_like(v[..., 0])
    rows = [
        [z, -v[..., 2], v[..., 1], z, z, z],
        [v[..., 2], z, -v[..., 0], z, z, z],
        [-v[..., 1], v[..., 0], z, z, z, z],
        [z, -v[..., 5], v[..., 4], z, -v[..., 2], v[..., 1]],
        [v[..., 5], z, -v[..., 3], v[..., 2], z, -v[..., 0]],
        [-v[..., 4], v[..., 3], z, -v[..., 1], v[..., 0], z],
    ]
    return np.stack([np.stack(r, axis=-1) for r in rows], axis=-2)


def mxS(S, vec, alpha=1.0):
    """alpha * (vec x) S   (RBDReference.py:58-63)."""
    a = np.asarray(alpha, dtype=np.float64)
    return a[..., None] * _mv(crm(vec), np.broadcast_to(S, vec.shape))


def fxv(m, f):
    """Force cross product  m x* f  (RBDReference.py:72-92)."""
    r = np.empty(np.broadcast_shapes(m.shape, f.shape))
    r[..., 0] = -m[..., 2] * f[..., 1] + m[..., 1] * f[..., 2] - m[..., 5] * f[..., 4] + m[..., 4] * f[..., 5]
    r[..., 1] = m[..., 2] * f[..., 0] - m[..., 0] * f[..., 2] + m[..., 5] * f[..., 3] - m[..., 3] * f[..., 5]
    r[..., 2] = -m[..., 1] * f[..., 0] + m[..., 0] * f[..., 1] - m[..., 4] * f[..., 3] + m[..., 3] * f[..., 4]
    r[..., 3] = -m[..., 2] * f[..., 4] + m[..., 1] * f[..., 5]
    r[..., 4] = m[..., 2] * f[..., 3] - m[..., 0] * f[..., 5]
    r[..., 5] = -m[..., 1] * f[..., 3] + m[..., 0] * f[..., 4]
    return r


def fxS(S, vec, alpha=1.0):
    """RBDReference.py:94-97."""
    return -mxS(S, vec, alpha)


def rnea(model, q, qd, qdd=None, gravity=-9.81):
    """RBDReference.rnea (:534) = rnea_fpass (:399-484) + rnea_bpass (:486-532).

    Returns c (..,n) and v, a, f (..,n,6); f is the *accumulated* joint force (as the reference returns it)."""
    q = np.asarray(q, dtype=np.float64)
    qd = np.asarray(qd, dtype=np.float64)
    n = model.n
    bshape = q.shape[:-1]
    v = np.zeros(bshape + (n, 6))
    a = np.zeros(bshape + (n, 6))
    f = np.zeros(bshape + (n, 6))
    gvec = np.zeros(6)
    gvec[5] = -gravity
    Xs = [model.X(j, q[..., j]) for j in range(n)]
    for j in range(n):
        p = model.parent[j]
        S = model.S[j]
        if p == -1:
            a[..., j, :] = _mv(Xs[j], np.broadcast_to(gvec, bshape + (6,)))
        else:
            v[..., j, :] = _mv(Xs[j], v[..., p, :])
            a[..., j, :] = _mv(Xs[j], a[..., p, :])
        v[..., j, :] += S * qd[..., j, None]
        a[..., j, :] += mxS(S, v[..., j, :], qd[..., j])
        if qdd is not None:
            a[..., j, :] += S * np.asarray(qdd)[..., j, None]
        Iv = _mv(model.I[j], v[..., j, :])
        f[..., j, :] = _mv(model.I[j], a[..., j, :]) + fxv(v[..., j, :], Iv)   # vxIv (:99-116)
    c = np.zeros(bshape + (n,))
    for j in range(n - 1, -1, -1):
        c[..., j] = np.sum(model.S[j] * f[..., j, :], axis=-1)
        p = model.parent[j]
        if p != -1:
            f[..., p, :] += _mtv(Xs[j], f[..., j, :])
    return c, v, a, f


def minv(model, q):
    """Analytical inverse of the joint-space inertia matrix.  RBDReference.minv (:908-930),
    minv_bpass (:805-876), minv_fpass (:878-906).  Returns (..,n,n), symmetric-filled from the upper triangle."""
    q = np.asarray(q, dtype=np.float64)
    n = model.n
    bshape = q.shape[:-1]
    Minv = np.zeros(bshape + (n, n))
    F = np.zeros(bshape + (n, 6, n))
    U = np.zeros(bshape + (n, 6))
    Dinv = np.zeros(bshape + (n,))
    IA = [np.broadcast_to(model.I[j], bshape + (6, 6)).copy() for j in range(n)]
    Xs = [model.X(j, q[..., j]) for j in range(n)]
    for j in range(n - 1, -1, -1):
        S = model.S[j]
        U[..., j, :] = _mv(IA[j], np.broadcast_to(S, bshape + (6,)))
        Dinv[..., j] = 1.0 / np.sum(S * U[..., j, :], axis=-1)
        Minv[..., j, j] = Dinv[..., j]
        for s in model.subtree[j]:
            Minv[..., j, s] -= Dinv[..., j] * np.sum(S * F[..., j, :, s], axis=-1)
        p = model.parent[j]
        if p != -1:
            for s in model.subtree[j]:
                F[..., j, :, s] += U[..., j, :] * Minv[..., j, s, None]
                F[..., p, :, s] += _mtv(Xs[j], F[..., j, :, s])
            Ia = IA[j] - U[..., j, :, None] * (Dinv[..., j, None] * U[..., j, :])[..., None, :]
            IA[p] = IA[p] + np.matmul(np.swapaxes(Xs[j], -1, -2), np.matmul(Ia, Xs[j]))
    for j in range(n):
        p = model.parent[j]
        S = model.S[j]
        if p != -1:
            UX = _mtv(Xs[j], U[..., j, :])                       # U^T X
            Minv[..., j, j:] -= Dinv[..., j, None] * np.matmul(UX[..., None, :], F[..., p, :, j:])[..., 0, :]
        F[..., j, :, j:] = S[:, None] * Minv[..., j, None, j:]
        if p != -1:
            F[..., j, :, j:] += np.matmul(Xs[j], F[..., p, :, j:])
    iu = np.triu_indices(n, 1)
    Minv[..., iu[1], iu[0]] = Minv[..., iu[0], iu[1]]
    return Minv


def rnea_grad(model, q, qd, qdd, gravity=-9.81):
    """d(rnea)/d(q,qd) -> (..,n,2n).  RBDReference.rnea_grad (:774-802) with the four passes
    rnea_grad_fpass_dq (:561-632), _fpass_dqd (:634-695), _bpass_dq (:697-734), _bpass_dqd (:736-772)."""
    q = np.asarray(q, dtype=np.float64)
    qd = np.asarray(qd, dtype=np.float64)
    n = model.n
    bshape = q.shape[:-1]
    c, v, a, f = rnea(model, q, qd, qdd, gravity)
    gvec = np.zeros(6)
    gvec[5] = -gravity
    Xs = [model.X(j, q[..., j]) for j in range(n)]
    # arrays [.., joint, column, 6]
    dv_dq = np.zeros(bshape + (n, n, 6)); da_dq = np.zeros(bshape + (n, n, 6)); df_dq = np.zeros(bshape + (n, n, 6))
    dv_dqd = np.zeros(bshape + (n, n, 6)); da_dqd = np.zeros(bshape + (n, n, 6)); df_dqd = np.zeros(bshape + (n, n, 6))
    for j in range(n):
        p = model.parent[j]
        S = model.S[j]
        X = Xs[j]
        I = model.I[j]
        Xc = X[..., None, :, :]
        # --- dq
        if p != -1:
            dv_dq[..., j, :, :] = _mv(Xc, dv_dq[..., p, :, :])
            dv_dq[..., j, j, :] += mxS(S, _mv(X, v[..., p, :]))
            da_dq[..., j, :, :] = _mv(Xc, da_dq[..., p, :, :])
        da_dq[..., j, :, :] += mxS(S, dv_dq[..., j, :, :], qd[..., j, None])
        if p != -1:
            da_dq[..., j, j, :] += mxS(S, _mv(X, a[..., p, :]))
        else:
            da_dq[..., j, j, :] += mxS(S, _mv(X, np.broadcast_to(gvec, bshape + (6,))))
        Iv = _mv(I, v[..., j, :])
        df_dq[..., j, :, :] = _mv(I, da_dq[..., j, :, :])
        df_dq[..., j, :, :] += fxv(dv_dq[..., j, :, :], Iv[..., None, :])
        df_dq[..., j, :, :] += fxv(v[..., j, None, :], _mv(I, dv_dq[..., j, :, :]))
        # --- dqd
        if p != -1:
            dv_dqd[..., j, :, :] = _mv(Xc, dv_dqd[..., p, :, :])
            da_dqd[..., j, :, :] = _mv(Xc, da_dqd[..., p, :, :])
        dv_dqd[..., j, j, :] += S
        da_dqd[..., j, :, :] += mxS(S, dv_dqd[..., j, :, :], qd[..., j, None])
        da_dqd[..., j, j, :] += mxS(S, v[..., j, :])
        df_dqd[..., j, :, :] = _mv(I, da_dqd[..., j, :, :])
        df_dqd[..., j, :, :] += fxv(dv_dqd[..., j, :, :], Iv[..., None, :])
        df_dqd[..., j, :, :] += fxv(v[..., j, None, :], _mv(I, dv_dqd[..., j, :, :]))
    dc_dq = np.zeros(bshape + (n, n))
    dc_dqd = np.zeros(bshape + (n, n))
    for j in range(n - 1, -1, -1):
        S = model.S[j]
        dc_dq[..., j, :] = np.sum(S * df_dq[..., j, :, :], axis=-1)
        dc_dqd[..., j, :] = np.sum(S * df_dqd[..., j, :, :], axis=-1)
        p = model.parent[j]
        if p != -1:
            XT = np.swapaxes(Xs[j], -1, -2)[..., None, :, :]
            df_dq[..., p, :, :] += _mv(XT, df_dq[..., j, :, :])
            df_dq[..., p, j, :] += _mtv(Xs[j], fxS(S, f[..., j, :]))
            df_dqd[..., p, :, :] += _mv(XT, df_dqd[..., j, :, :])
    return np.concatenate([dc_dq, dc_dqd], axis=-1)


# ---------------------------------------------------------------------------------------------------
# planar end-effector kinematics (RBDReference.py:123-148, 219-266, 318-387); serial chains only
# ---------------------------------------------------------------------------------------------------
EE_OFFSET = np.array([0.0, 1.0, 0.0, 1.0])


def _chain(model):
    leaf = model.leaves()[0]
    return sorted(model.ancestors(leaf)) + [leaf]


def end_effector_positions(model, q, offset=EE_OFFSET):
    """(x, y) of the first leaf's end effector, (..,2).  RBDReference.py:123-148."""
    q = np.asarray(q, dtype=np.float64)
    T = np.broadcast_to(np.eye(4), q.shape[:-1] + (4, 4))
    for j in _chain(model):
        T = np.matmul(T, model.H(j, q[..., j]))
    return _mv(T, np.broadcast_to(offset, q.shape[:-1] + (4,)))[..., :2]


def jacobian(model, q, offset=EE_OFFSET):
    """RBDReference.Jacobian (:339-387): rows = first min(3,n) of (x,y,z), cols = joints -> (..,n',n) with n'=min(3,n)."""
    q = np.asarray(q, dtype=np.float64)
    n = model.n
    chain = _chain(model)
    cols = []
    for d in range(n):
        if d not in chain:
            cols.append(np.zeros(q.shape[:-1] + (3,)))
            continue
        T = np.broadcast_to(np.eye(4), q.shape[:-1] + (4, 4))
        for j in chain:
            T = np.matmul(T, model.dH(j, q[..., j]) if j == d else model.H(j, q[..., j]))
        cols.append(_mv(T, np.broadcast_to(offset, q.shape[:-1] + (4,)))[..., :3])
    J = np.stack(cols, axis=-1)          # (..,3,n)
    return J[..., :n, :n]


def dJdq(model, q, offset=EE_OFFSET):
    """RBDReference.dJdq (:219-266) -- literal restatement, valid for n == 2 only (the reference hard-codes 2 columns)."""
    assert model.n == 2, "the reference's end-effector Jacobian derivative is hard-wired to 2 joints (RBDReference.py:263)"
    J = jacobian(model, q, offset)
    out = np.zeros(q.shape[:-1] + (4, 2))
    out[..., 0, :] = -J[..., 1, :]
    out[..., 1, 0] = -J[..., 1, 1]
    out[..., 1, 1] = -J[..., 1, 1]
    out[..., 2, :] = -J[..., 0, :]
    out[..., 3, 0] = J[..., 0, 1]
    out[..., 3, 1] = J[..., 0, 1]
    return out


def planar_jacobians(model, q, offset=EE_OFFSET):
    """Exact first and second derivatives of the planar end-effector position for an n-joint chain (UNPINNED, SURVEY.md 8f-3:
    the intent of RBDReference_generalized.py:425-459).  Returns J (..,2,n) = d(x,y)/dq and Hs (..,2,n,n) = d2(x,y)/dq_c dq_d."""
    q = np.asarray(q, dtype=np.float64)
    n = model.n
    chain = _chain(model)
    off = np.broadcast_to(offset, q.shape[:-1] + (4,))

    def prod(orders):
        T = np.broadcast_to(np.eye(4), q.shape[:-1] + (4, 4))
        for j in chain:
            o = orders.get(j, 0)
            M = model.H(j, q[..., j]) if o == 0 else (model.dH(j, q[..., j]) if o == 1 else model.ddH(j, q[..., j]))
            T = np.matmul(T, M)
        return _mv(T, off)[..., :2]

    J = np.zeros(q.shape[:-1] + (2, n))
    Hs = np.zeros(q.shape[:-1] + (2, n, n))
    for d in chain:
        J[..., :, d] = prod({d: 1})
        for c in chain:
            if c > d:
                continue
            v = prod({d: 2}) if c == d else prod({c: 1, d: 1})
            Hs[..., :, c, d] = v
            Hs[..., :, d, c] = v
    return J, Hs


def planar_third_derivative_times_qd(model, q, qd, offset=EE_OFFSET):
    """sum_j d3(x,y)/dq_a dq_b dq_j qd_j  -> (..,2,n,n): the q-q block of the second derivative of the end-effector velocity J(q) qd
    (exact; needed by the exact Hessian of the end-effector cost, UrdfCost.hess_mode 1)."""
    q = np.asarray(q, dtype=np.float64)
    qd = np.asarray(qd, dtype=np.float64)
    n = model.n
    chain = _chain(model)
    off = np.broadcast_to(offset, q.shape[:-1] + (4,))
    mats = {0: model.H, 1: model.dH, 2: model.ddH, 3: model.dddH}

    def prod(orders):
        T = np.broadcast_to(np.eye(4), q.shape[:-1] + (4, 4))
        for j in chain:
            T = np.matmul(T, mats[orders.get(j, 0)](j, q[..., j]))
        return _mv(T, off)[..., :2]

    out = np.zeros(q.shape[:-1] + (2, n, n))
    for ia, a in enumerate(chain):
        for b in chain[ia:]:
            acc = np.zeros(q.shape[:-1] + (2,))
            for j in chain:
                orders = {}
                for idx in (a, b, j):
                    orders[idx] = orders.get(idx, 0) + 1
                acc = acc + prod(orders) * qd[..., j, None]
            out[..., :, a, b] = acc
            out[..., :, b, a] = acc
    return out


def jacobian_tot_state_general(model, q, qd, offset=EE_OFFSET):
    """d(x,y,vx,vy)/d(q,qd) for an n-joint planar chain, exact:  [[J, 0], [d(J qd)/dq, J]]  -> (..,4,2n)."""
    J, Hs = planar_jacobians(model, q, offset)
    J2 = np.einsum("...rcd,...c->...rd", Hs, np.asarray(qd, dtype=np.float64))
    top = np.concatenate([J, np.zeros_like(J)], axis=-1)
    bot = np.concatenate([J2, J], axis=-1)
    return np.concatenate([top, bot], axis=-2)


def jacobian_tot_state(model, q, qd, offset=EE_OFFSET):
    """d(x,y,vx,vy)/d(q,qd) as the reference builds it (RBDReference.py:318-336) -> (..,2n,2n), n == 2."""
    n = model.n
    J1 = jacobian(model, q, offset)
    D = dJdq(model, q, offset)
    J2 = _mv(D, np.asarray(qd, dtype=np.float64)).reshape(q.shape[:-1] + (n, n))
    top = np.concatenate([J1, np.zeros_like(J1)], axis=-1)
    bot = np.concatenate([J2, J1], axis=-1)
    return np.concatenate([top, bot], axis=-2)
