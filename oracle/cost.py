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
