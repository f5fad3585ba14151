"""Soft box-limit oracle (QUADRATIC_PENALTY / AUGMENTED_LAGRANGIAN), element-wise restatement.

Restates /root/reference/TrajoptConstraint.py: BoxConstraint.value (:53-90), .jacobian (:92-129),
.max_soft_constraint_value (:131-136), .update_soft_constraint_constants (:138-166), and the TrajoptConstraint
aggregation (:295-378).  The reference's BoxConstraint only works for constraint_size == 1;
for one constrained coordinate this restatement is identical to it (pinned by tests/golden/solve_pend_*.npz),
for several coordinates it applies the same formulas per coordinate (UNPINNED -- this file is the spec):

    v       = [z - lb ; ub - z]                                  (2 cs)
    value_k = sum_i mu_ik v_i^2   (+ sum_i lambda_ik v_i   if AL)          counts inactive sides too
    gck     = sum_{i: v_i<0} (2 mu_ik v_i (+ lambda_ik if AL)) grad(v_i)   grad(v_i) = +e_i (lower), -e_i (upper)
    QP terms: g_k += gck ; G_k += gck gck^T                      (TrajoptMPCReference.py:220-224)

Deviations from the (crashing) reference for the unpinned multi-limit case, stated here once:
  * joint limits carry N columns of (mu, lambda, phi) (the reference allocates N-1 and indexes column N-1),
  * velocity limits act on x[nq:nq+nv] (the reference's value() reads x[:nv]),
  * several limit types add their gradients (the reference vstacks (m,1) columns, a shape error),
  * every limit type is updated in the outer loop (the reference's `flag and update()` short-circuits).
"""
import numpy as np

SOFT_MODES = ("QUADRATIC_PENALTY", "AUGMENTED_LAGRANGIAN")
HARD_MODES = ("ACTIVE_SET",)        # FULL_SET adds the inactive rows too (zero Jacobian rows): singular KKT in the reference


class BoxLimit:
    def __init__(self, offset, size, T, upper, lower, mode, options=None):
        if mode not in SOFT_MODES + HARD_MODES:
            raise ValueError("oracle supports %r" % (SOFT_MODES + HARD_MODES,))
        self.hard = mode in HARD_MODES
        o = dict(options or {})
        self.mu_init = o.get("quadratic_penalty_mu_init", 1e-2)
        self.mu_factor = o.get("quadratic_penalty_mu_factor", 10.0)
        self.mu_max = o.get("quadratic_penalty_mu_max", 1e12)
        self.phi_init = o.get("augmentated_lagrangian_phi_init", 1e-2)
        self.phi_factor = o.get("augmentated_lagrangian_phi_factor", 10.0)
        self.offset, self.size, self.T, self.mode = offset, size, T, mode
        ub = np.broadcast_to(np.asarray(upper, dtype=np.float64), (size,))
        lb = np.broadcast_to(np.asarray(lower, dtype=np.float64), (size,))
        self.lb, self.ub = lb.copy(), ub.copy()
        self.mu = self.mu_init * np.ones((2 * size, T))
        self.lam = np.zeros((2 * size, T))
        self.phi = self.phi_init * np.ones((2 * size, T))

    def v(self, Z):
        """Z (T, m) knot-major stacked [x;u] rows -> (T, 2cs)"""
        z = Z[:, self.offset:self.offset + self.size]
        return np.concatenate([z - self.lb, self.ub - z], axis=-1)

    def values(self, Z):
        v = self.v(Z)
        T = v.shape[0]
        val = np.einsum("ik,ki->k", self.mu[:, :T], v * v)
        if self.mode == "AUGMENTED_LAGRANGIAN":
            val = val + np.einsum("ik,ki->k", self.lam[:, :T], v)
        return val

    def gradients(self, Z):
        """(T, m) summed penalty gradient gck"""
        v = self.v(Z)
        T = v.shape[0]
        act = v < 0
        coef = 2.0 * self.mu[:, :T].T * v
        if self.mode == "AUGMENTED_LAGRANGIAN":
            coef = coef + self.lam[:, :T].T
        coef = np.where(act, coef, 0.0)
        g = np.zeros_like(Z)
        cs = self.size
        g[:, self.offset:self.offset + cs] = coef[:, :cs] - coef[:, cs:]
        return g

    def max_value(self, Z):
        """max over knots of |min_i v_i|  (literal: positive margins count too, :131-136)"""
        v = self.v(Z)
        return float(np.max(np.abs(np.min(v, axis=1)))) if v.size else 0.0

    def update(self, Z):
        """mu / lambda / phi update (:138-166); returns the 'all mu at max (or nothing active)' flag."""
        v = self.v(Z)
        T = v.shape[0]
        flag = True
        for k in range(T):
            for i in range(2 * self.size):
                if not (v[k, i] < 0):
                    continue
                if not (abs(v[k, i]) < self.phi[i, k]):
                    if self.mu[i, k] < self.mu_max:
                        flag = False
                        self.mu[i, k] = min(self.mu_max, self.mu[i, k] * self.mu_factor)
                else:
                    flag = False
                    self.lam[i, k] += self.mu[i, k] * v[k, i]
                    self.phi[i, k] /= self.phi_factor
        return flag


class SoftConstraints:
    """TrajoptConstraint restricted to soft box limits (TrajoptConstraint.py:178-208, 295-378)."""

    def __init__(self, nq, nv, nu, N):
        self.nq, self.nv, self.nu, self.N = nq, nv, nu, N
        self.limits = {}

    def set_joint_limits(self, upper, lower, mode, options=None):
        self.limits["joint"] = BoxLimit(0, self.nq, self.N, upper, lower, mode, options)

    def set_velocity_limits(self, upper, lower, mode, options=None):
        self.limits["velocity"] = BoxLimit(self.nq, self.nv, self.N, upper, lower, mode, options)

    def set_torque_limits(self, upper, lower, mode, options=None):
        self.limits["torque"] = BoxLimit(self.nq + self.nv, self.nu, self.N - 1, upper, lower, mode, options)

    def any(self):
        """any SOFT limit (total_soft_constraints() > 0, TrajoptConstraint.py:281-293)"""
        return any(not lim.hard for lim in self.limits.values())

    def any_hard(self):
        return any(lim.hard for lim in self.limits.values())

    def hard_rows(self, X, U):
        """ACTIVE_SET hard constraints (BoxConstraint.value / .jacobian with mode ACTIVE_SET, TrajoptConstraint.py:53-67, 92-112;
        aggregation value_hard_constraints / jacobian_hard_constraints :210-279): per knot the violated bounds only, as
        (rows (r, m), values (r,)) in the reference's order -- limit types joint, velocity, torque; inside a type all violated lower
        bounds, then all violated upper bounds.  value = z - lb (lower) resp. ub - z (upper), both < 0; Jacobian row +e_i resp. -e_i."""
        Z = self._Z(X, U)
        N, m = Z.shape
        out = []
        for k in range(N):
            rows, vals = [], []
            for name in ("joint", "velocity", "torque"):
                lim = self.limits.get(name)
                if lim is None or not lim.hard or k >= lim.T:
                    continue
                v = lim.v(Z[k:k + 1])[0]
                for i in range(2 * lim.size):
                    if v[i] < 0:
                        row = np.zeros(m)
                        row[lim.offset + (i % lim.size)] = 1.0 if i < lim.size else -1.0
                        rows.append(row); vals.append(v[i])
            out.append((np.array(rows).reshape(len(vals), m), np.array(vals)))
        return out

    def _Z(self, X, U):
        N = X.shape[0]
        Z = np.zeros((N, self.nq + self.nv + self.nu))
        Z[:, :self.nq + self.nv] = X
        Z[:N - 1, self.nq + self.nv:] = U
        return Z

    def values(self, X, U):
        Z = self._Z(X, U)
        val = np.zeros(X.shape[0])
        for name, lim in self.limits.items():
            if lim.hard:
                continue
            T = lim.T
            val[:T] += lim.values(Z[:T])
        return val

    def gradients(self, X, U):
        Z = self._Z(X, U)
        g = np.zeros_like(Z)
        for name, lim in self.limits.items():
            if lim.hard:
                continue
            T = lim.T
            g[:T] += lim.gradients(Z[:T])
        return g

    def max_value(self, X, U):
        Z = self._Z(X, U)
        mx = 0.0
        for lim in self.limits.values():
            if not lim.hard:
                mx = max(mx, lim.max_value(Z[:lim.T]))
        return mx

    def update(self, X, U):
        Z = self._Z(X, U)
        flag = True
        for lim in self.limits.values():
            if lim.hard:
                continue
            f = lim.update(Z[:lim.T])
            flag = flag and f
        return flag
