"""Cost oracles: QuadraticCost (joint space) and UrdfCost (planar end-effector space, n == 2).

Restates /root/reference/TrajoptCost.py:24-104 (QuadraticCost) and :371-519 (UrdfCost, hess_mode 0).
Trajectory layout: X (N, nx) rows = knot points, U (N-1, nu).  The last knot has no control (terminal cost, QF).
"""
import numpy as np

from . import rbd


class QuadraticCost:
    def __init__(self, Q, QF, R, xg, QF_start=None):
        self.Q = np.asarray(Q, dtype=np.float64)
        self.QF = np.asarray(QF, dtype=np.float64)
        self.R = np.asarray(R, dtype=np.float64)
        self.xg = np.asarray(xg, dtype=np.float64)
        self.QF_start = QF_start
        self.nx = self.Q.shape[0]
        self.nu = self.R.shape[0]

    def currQ(self, k, terminal):
        """get_currQ (TrajoptCost.py:40-47): QF at the terminal knot or from QF_start on."""
        use_qf = terminal or (self.QF_start is not None and k >= self.QF_start)
        return self.QF if use_qf else self.Q

    def _Qs(self, N):
        return np.stack([self.currQ(k, k == N - 1) for k in range(N)])

    def state_error(self, X):
        return X - self.xg

    def state_jacobian(self, X):
        return None      # identity

    def values(self, X, U):
        """per-knot cost (N,)  (value :49-56)"""
        N = X.shape[0]
        dx = self.state_error(X)
        Qs = self._Qs(N)
        val = 0.5 * np.einsum("ki,kij,kj->k", dx, Qs, dx)
        val[:N - 1] += 0.5 * np.einsum("ki,ij,kj->k", U, self.R, U)
        return val

    def gradients(self, X, U):
        """(N, nx+nu); the terminal row's control part is zero  (gradient :58-69)"""
        N = X.shape[0]
        dx = self.state_error(X)
        Qs = self._Qs(N)
        g = np.zeros((N, self.nx + self.nu))
        top = np.einsum("ki,kij->kj", dx, Qs)               # delta_x^T Q
        Jt = self.state_jacobian(X)
        g[:, :self.nx] = top if Jt is None else np.einsum("ki,kij->kj", top, Jt)
        g[:N - 1, self.nx:] = np.einsum("ki,ij->kj", U, self.R)
        return g

    def hessians(self, X, U):
        """(N, m, m) block-diag(Q_k, R); terminal block only [:nx,:nx]  (hessian :71-83)"""
        N = X.shape[0]
        m = self.nx + self.nu
        H = np.zeros((N, m, m))
        Qs = self._Qs(N)
        Jt = self.state_jacobian(X)
        if Jt is None:
            H[:, :self.nx, :self.nx] = Qs
        else:
            QJ = np.matmul(Qs, Jt)
            H[:, :self.nx, :self.nx] = np.matmul(np.swapaxes(QJ, -1, -2), Jt)    # ((Q J)^T) J  (:493)
            if getattr(self, "hess_mode", 0) == 1:
                H[:, :self.nx, :self.nx] += self.second_order_term(X, Qs)
        H[:N - 1, self.nx:, self.nx:] = self.R
        return H


class UrdfCost(QuadraticCost):
    """Quadratic cost on the end-effector state (x, y, vx, vy); Gauss-Newton Hessian (hess_mode 0).
    TrajoptCost.py:371-519.  n == 2 restates the reference literally (including its dJdq, RBDReference.py:256-266) and is pinned;
    n > 2 is the exact generalisation to an n-joint planar chain (UNPINNED; intent in
    TrajoptCost_generalized.py:405-467): Q, QF are 4x4, xg = (x, y, vx, vy), J_tot is 4 x 2n."""

    def __init__(self, model, Q, QF, R, xg, QF_start=None):
        super().__init__(Q, QF, R, xg, QF_start)
        self.model = model
        self.n = model.n
        self.nx = 2 * self.n
        assert self.n >= 2 and self.Q.shape == (4, 4)

    hess_mode = 0       # 0: Gauss-Newton J_tot^T Q J_tot (:493); 1: exact = Gauss-Newton + sum_i (Q e)_i d2 e_i / dx2

    def second_order_term(self, X, Qs):
        """hess_mode 1 (TrajoptCost.py:494-499 computes `hess_1 + (dx^T Q dJtotdq)` but crashes: `hess_x` is unbound and
        RBDReference has no d2Jdq2): the intent is the EXACT Hessian of 0.5 e^T Q e, i.e. the Gauss-Newton part plus
        sum_i (Q e)_i * Hessian(e_i), e = [p(q); J(q) qd] - xg.  UNPINNED: this exact form is the specification, for every n >= 2.
            d2 p_r / dq dq = Hs_r;   d2 (J qd)_r / dq dq = sum_j d3 p_r/dq dq dq_j qd_j;   d2 (J qd)_r / dq dqd = Hs_r."""
        n, nx = self.n, self.nx
        q, qd = X[..., :n], X[..., n:]
        e = self.state_error(X)
        w = np.matmul(Qs, e[..., None])[..., 0]                     # (N, 4)
        _, Hs = rbd.planar_jacobians(self.model, q)                  # (N, 2, n, n)
        T3 = rbd.planar_third_derivative_times_qd(self.model, q, qd)
        S2 = np.zeros(X.shape[:-1] + (nx, nx))
        for r in range(2):
            S2[..., :n, :n] += w[..., r, None, None] * Hs[..., r, :, :] + w[..., 2 + r, None, None] * T3[..., r, :, :]
            S2[..., :n, n:] += w[..., 2 + r, None, None] * Hs[..., r, :, :]
            S2[..., n:, :n] += w[..., 2 + r, None, None] * Hs[..., r, :, :]
        return S2

    def state_error(self, X):
        """delta_x (:425-435): [ee_pos; J qd] - xg"""
        n = self.n
        pos = rbd.end_effector_positions(self.model, X[..., :n])
        J = rbd.jacobian(self.model, X[..., :n])[..., :2, :] if n == 2 else rbd.planar_jacobians(self.model, X[..., :n])[0]
        vel = np.matmul(J, X[..., n:, None])[..., 0]
        return np.concatenate([pos, vel], axis=-1) - self.xg

    def state_jacobian(self, X):
        n = self.n
        if n == 2:
            return rbd.jacobian_tot_state(self.model, X[..., :n], X[..., n:])
        return rbd.jacobian_tot_state_general(self.model, X[..., :n], X[..., n:])
