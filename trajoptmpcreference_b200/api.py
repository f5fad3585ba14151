"""Drop-in Python API of the reference (TrajoptPlant / TrajoptCost / TrajoptConstraint / TrajoptMPCReference) on top of
the per-robot CUDA libraries.  Same class names, constructor signatures, option keys and return values as
/root/reference/{TrajoptPlant,TrajoptCost,TrajoptConstraint,TrajoptMPCReference}.py for the SQP-PCG path; everything
numerical happens on the GPU through the C ABI (include/b2t.h).  No CPU fallback.

Differences from the reference, all deliberate:
  * invalid arguments raise ValueError instead of print() + exit()   (TrajoptMPCReference.py:32-39, 596)
  * the solver is re-entrant: no class-level counters, no unbounded saved_* lists
  * costs are described by their parameters (Q, QF, R, xg); arbitrary Python callbacks cannot run inside a kernel
  * `solve_batch` solves many independent instances (different goals / initial trajectories) in one launch sequence
"""
import ctypes
import hashlib
import enum
import os

import numpy as np

from . import _lib
from .model import extract_model, model_digest, URDF_DIR


class SQPSolverMethods(enum.Enum):          # TrajoptMPCReference.py:13-18
    N = "N"
    S = "S"
    PCG_J = "PCG-J"
    PCG_BJ = "PCG-BJ"
    PCG_SS = "PCG-SS"


class MPCSolverMethods(enum.Enum):          # TrajoptMPCReference.py:21-27
    iLQR = "iLQR"
    QP_N = "QP-N"
    QP_S = "QP-S"
    QP_PCG_J = "QP-PCG-J"
    QP_PCG_BJ = "QP-PCG-BJ"
    QP_PCG_SS = "QP-PCG-SS"


_METHOD_CODE = {SQPSolverMethods.N: _lib.METHOD_N, SQPSolverMethods.S: _lib.METHOD_S, SQPSolverMethods.PCG_J: _lib.METHOD_PCG_J,
                SQPSolverMethods.PCG_BJ: _lib.METHOD_PCG_BJ, SQPSolverMethods.PCG_SS: _lib.METHOD_PCG_SS}


# --------------------------------------------------------------------------------------------------- plant
class _RbdHandle:
    """The solver and UrdfCost read / write `plant.rbdReference.overloading` (TrajoptMPCReference.py:97, TrajoptCost.py:389)."""

    def __init__(self, model):
        self.model = model
        self.overloading = False


class TrajoptPlant:
    def __init__(self, integrator_type: int = 0, options=None, need_path: bool = False):
        options = {} if options is None else options
        if integrator_type not in (0, 1, 2, 3, 4, -1):
            raise ValueError("Invalid integrator options are [0 : euler, 1 : semi-implicit euler, 2 : midpoint, 3 : rk3, 4 : rk4, -1 : hard-coded as dynamics")
        if integrator_type == -1:
            raise ValueError("integrator type -1 (dynamics hard-coded in a user plant class) cannot run inside the kernels; use 0-3")
        # 2 (midpoint) and 3 (rk3) reproduce the reference's own arithmetic, whose Jacobians do not match its step;
        # 4 (rk4) steps, but its gradient -- hence every solve -- raises TypeError exactly like TrajoptPlant.py:259
        self.integrator_type = integrator_type
        options.setdefault("path_to_urdf", None)
        options.setdefault("gravity", -9.81)
        if need_path and not options.get("path_to_urdf"):
            raise ValueError("You must include the 'path_to_urdf' in the options.")
        self.options = options

    def forward_dynamics(self, *a, **k):
        raise NotImplementedError

    def forward_dynamics_gradient(self, *a, **k):
        raise NotImplementedError

    def get_num_pos(self):
        raise NotImplementedError

    def get_num_vel(self):
        raise NotImplementedError

    def get_num_cntrl(self):
        raise NotImplementedError


class URDFPlant(TrajoptPlant):
    """URDFPlant (TrajoptPlant.py:274-331).  `options['path_to_urdf']` may also be a built-in name ('arm6', 'pend', ...)."""

    def __init__(self, integrator_type=0, options=None):
        options = {} if options is None else options
        super().__init__(integrator_type, options, True)
        path = options["path_to_urdf"]
        if not os.path.isfile(path):
            cand = os.path.join(URDF_DIR, path + ".urdf")
            if os.path.isfile(cand):
                path = cand
            else:
                raise ValueError("Failed to parse URDF file at the given path.")
        self.urdf_path = path
        self.model = extract_model(path)
        base = os.path.splitext(os.path.basename(path))[0]
        builtin = os.path.join(URDF_DIR, base + ".urdf")
        same = os.path.isfile(builtin) and model_digest(extract_model(builtin)) == model_digest(self.model)
        self.tag = base if same else "%s_%s" % ("".join(ch if ch.isalnum() else "_" for ch in base), model_digest(self.model))
        self.rbdReference = _RbdHandle(self.model)
        self._lib = None
        self._probe = None
        # plant-level recording lists of the reference (TrajoptPlant.py:297-299, 318-322): one entry per callback invocation, filled
        # while `recording` is on (SQP(record=True) switches it on and replays the reference's call sequence, record.py)
        self.recording = False
        self.saved_c, self.saved_Minv, self.saved_qdd, self.saved_dc_du, self.saved_dqdd = [], [], [], [], []

    @property
    def lib(self):
        if self._lib is None:
            self._lib = _lib.load_library(self.model, self.tag)
        return self._lib

    def get_num_pos(self):
        return self.model["n"]

    def get_num_vel(self):
        return self.model["n"]

    def get_num_cntrl(self):
        return self.model["n"]

    # per-knot callbacks of the reference API, evaluated on the GPU (batch 1, 2 knots)
    def _probe_solver(self, dt=None, own_integrator=False):
        """Batch-1, 2-knot workspace for the per-knot callbacks.  The default probe integrates with Euler (dqdd, qdd, Minv, ... do not
        depend on the integrator); `own_integrator` selects the plant's multi-stage integrator (types 2 / 3) for x+ and (A, B)."""
        slot = "_probe_own" if own_integrator else "_probe"
        p = getattr(self, slot, None)
        if p is None or (dt is not None and p.dt != dt):
            n = self.model["n"]
            cost = QuadraticCost(np.eye(2 * n), np.eye(2 * n), np.eye(n), np.zeros(2 * n))
            it = self.integrator_type if (own_integrator or self.integrator_type in (0, 1)) else 0
            p = BatchSolver(self, cost, None, N=2, dt=1.0 if dt is None else dt, batch=1, integrator_type=it)
            setattr(self, slot, p)
        return p

    def _load_point(self, p, xk, uk):
        n = self.model["n"]
        x = np.zeros((1, 2 * n, 2)); u = np.zeros((1, n, 1))
        x[0, :, 0] = np.asarray(xk, dtype=np.float64).reshape(-1); u[0, :, 0] = np.asarray(uk, dtype=np.float64).reshape(-1)
        p.set_trajectory(x, u)

    def _eval(self, xk, uk, dt):
        p = self._probe_solver(dt)
        n = self.model["n"]
        self._load_point(p, xk, uk)
        p.stage_dynamics()
        return p.fetch("dqdd")[0, 0].reshape(n, 3 * n), p.fetch("xkp1")[0, 0]

    def _terms(self, xk, uk):
        """(c, qdd, Minv, dc_du) at one knot from the plant kernel (B2T_ARR_PLANT_TERMS)."""
        p = self._probe_solver()
        n = self.model["n"]
        self._load_point(p, xk, uk)
        return split_plant_terms(p.fetch("plant_terms")[0, 0], n)

    def _save(self, tag, **entries):
        for name, value in entries.items():
            getattr(self, "saved_" + name).append(dict(value=value, **tag))

    def forward_dynamics_gradient(self, x, u, iter_1=0, iter_2=0, iter_3=0):
        dqdd = self._eval(x, u, 1.0)[0]
        if self.recording:
            c, qdd, Minv, dc_du = self._terms(x, u)
            self._save(dict(iteration=iter_1, outer_iteration=iter_2, line_search_iteration=iter_3), Minv=Minv, c=c, qdd=qdd, dc_du=dc_du, dqdd=dqdd)
        return dqdd

    def forward_dynamics(self, x, u, iter_1=0, iter_2=0, iter_3=0):
        c, qdd, Minv, _ = self._terms(x, u)
        if self.recording:
            self._save(dict(iteration=iter_1, outer_iteration=iter_2, line_search_iteration=iter_3), c=c, Minv=Minv, qdd=qdd)
        return qdd

    def integrator(self, xk, uk, dt, return_gradient=False, iter_1=0, iter_2=0, iter_3=0):
        n = self.model["n"]
        if self.integrator_type in (2, 3):
            # midpoint / rk3 with the reference's own arithmetic (TrajoptPlant.py:140-205): x+ from k_fd, [A B] from k_ab_multi
            p = self._probe_solver(dt, own_integrator=True)
            self._load_point(p, xk, uk)
            p.stage_dynamics()
            if not return_gradient:
                return p.fetch("xkp1")[0, 0]
            AB = p.fetch("AB")[0, 0].reshape(2 * n, 3 * n)
            return AB[:, :2 * n].copy(), AB[:, 2 * n:].copy()
        if self.integrator_type == 4:
            if return_gradient:      # the reference's numpy branch passes an extra positional xk here (TrajoptPlant.py:259)
                raise TypeError("URDFPlant.forward_dynamics_gradient() takes from 3 to 6 positional arguments but 7 were given "
                                "(integrator type 4 of the reference, TrajoptPlant.py:259)")
            xk = np.asarray(xk, dtype=np.float64).reshape(-1)
            xdot = lambda pt: np.concatenate([xk[n:], self.forward_dynamics(pt, uk, iter_1, iter_2, iter_3)])      # noqa: E731  (:61-70: x_k's velocity)
            xdot1 = xdot(xk)
            xdot2 = xdot(xk + 0.5 * dt * xdot1)
            xdot3 = xdot(xk + 0.5 * dt * xdot2)
            xdot4 = xdot(xk + dt * xdot3)
            return xk + (dt / 6) * (xdot1 + 2 * xdot2 + 2 * xdot3 + xdot4)
        dqdd, xn = self._eval(xk, uk, dt)
        if not return_gradient:
            return xn
        top = np.hstack((np.zeros((n, n)), np.eye(n), np.zeros((n, n))))
        if self.integrator_type == 0:
            dxdot = np.vstack((top, dqdd))
            return np.eye(2 * n) + dt * dxdot[:, :2 * n], dt * dxdot[:, 2 * n:]
        Iz = np.hstack((np.eye(2 * n), np.zeros((2 * n, n))))
        AB = Iz + dt * np.vstack((top + dt * dqdd, dqdd))
        return AB[:, :2 * n], AB[:, 2 * n:]


def split_plant_terms(t, n):
    """[c | qdd | Minv | d rnea / d(q, qd)] of one knot (B2T_ARR_PLANT_TERMS) -> (c (n,), qdd (n,), Minv (n, n), dc_du (n, 2n))."""
    return t[:n].copy(), t[n:2 * n].copy(), t[2 * n:2 * n + n * n].reshape(n, n).copy(), t[2 * n + n * n:].reshape(n, 2 * n).copy()


# --------------------------------------------------------------------------------------------------- costs
class TrajoptCost:
    def value(self, *a, **k):
        raise NotImplementedError

    def gradient(self, *a, **k):
        raise NotImplementedError

    def hessian(self, *a, **k):
        raise NotImplementedError


class QuadraticCost(TrajoptCost):
    """QuadraticCost (TrajoptCost.py:24-104); value/gradient/hessian accept the 3-argument form and the iter_1..3 kwargs."""
    _kind = _lib.COST_QUADRATIC

    def __init__(self, Q_in, QF_in, R_in, xg_in, QF_start=None):
        self.Q = Q_in
        self.QF = QF_in
        self.R = R_in
        self.xg = xg_in
        self.increaseCount_Q = 0
        self.increaseCount_QF = 0
        self.QF_start = QF_start

    def get_currQ(self, u=None, timestep=None):
        use_QF = (u is None) or (timestep is not None and self.QF_start is not None and timestep >= self.QF_start)
        return self.QF if use_QF else self.Q

    def value(self, x, u=None, timestep=None, iter_1=0, iter_2=0, iter_3=0):
        dx = np.asarray(x) - self.xg
        cost = 0.5 * np.matmul(dx.transpose(), np.matmul(self.get_currQ(u, timestep), dx))
        if u is not None:
            cost += 0.5 * np.matmul(np.asarray(u).transpose(), np.matmul(self.R, u))
        return cost

    def gradient(self, x, u=None, timestep=None, iter_1=0, iter_2=0, iter_3=0):
        dx = np.asarray(x) - self.xg
        top = np.matmul(dx.transpose(), self.get_currQ(u, timestep))
        if u is None:
            return top
        return np.hstack((top, np.matmul(np.asarray(u).transpose(), self.R)))

    def hessian(self, x, u=None, timestep=None, iter_1=0, iter_2=0, iter_3=0):
        nx, nu = np.asarray(self.Q).shape[0], np.asarray(self.R).shape[0]
        currQ = self.get_currQ(u, timestep)
        if u is None:
            return currQ
        return np.vstack((np.hstack((currQ, np.zeros((nx, nu)))), np.hstack((np.zeros((nu, nx)), self.R))))

    def increase_QF(self, multiplier: float = 2.0):
        self.QF *= multiplier
        self.increaseCount_QF += 1
        return self.increaseCount_QF

    def increase_Q(self, multiplier: float = 2.0):
        self.Q *= multiplier
        self.increaseCount_Q += 1
        return self.increaseCount_Q

    def reset_increase_count_QF(self):
        self.increaseCount_QF = 0

    def reset_increase_count_Q(self):
        self.increaseCount_Q = 0

    def shift_QF_start(self, shift: float = -1.0):
        self.QF_start += shift
        self.QF_start = max(self.QF_start, 0)
        return self.QF_start


class UrdfCost(QuadraticCost):
    """UrdfCost (TrajoptCost.py:371-519): quadratic cost on the planar end-effector state (x, y, vx, vy), Gauss-Newton
    Hessian (hess_mode 0).  Two joints: the reference's arithmetic, literally.  More than two joints (a planar
    serial chain): the exact generalisation the reference only sketches (TrajoptCost_generalized.py:405-467):
    Q, QF stay 4 x 4, xg = (x, y, vx, vy), J_tot = [[J, 0], [d(J qd)/dq, J]] is 4 x 2n."""
    _kind = _lib.COST_URDF_EE

    def __init__(self, plant, Q_in, QF_in, R_in, xg_in, QF_start=None, overloading=False):
        super().__init__(Q_in, QF_in, R_in, xg_in, QF_start)
        if plant.get_num_pos() < 2:
            raise ValueError("UrdfCost needs a planar chain of at least 2 joints")
        if np.asarray(Q_in).shape != (4, 4) or np.asarray(QF_in).shape != (4, 4) or np.asarray(xg_in).size != 4:
            raise ValueError("UrdfCost weighs the end-effector state (x, y, vx, vy): Q, QF must be 4 x 4 and xg of length 4")
        self.plant = plant
        self.n = plant.get_num_pos()
        self.offsets = [np.array([[0, 1, 0, 1]])]
        self.plant.rbdReference.overloading = overloading
        self.overloading = overloading
        self.hess_mode = 0
        # cost-level recording lists of the reference (TrajoptCost.py:382-386, appended at :411-412, :420-421, :457-458, :516-517)
        self.recording = False
        self.saved_cost, self.saved_grad, self.saved_hess, self.saved_Jacobian_tot_state, self.saved_dx = [], [], [], [], []

    def _knot(self, x, u, timestep):
        """Evaluate this cost at one knot on the GPU (1 instance, 2 knots: knot 0 = (x,u) as a running knot, knot 1 = x as terminal)."""
        n = self.n
        use_qf_running = timestep is not None and self.QF_start is not None and timestep >= self.QF_start
        key = ("probe", use_qf_running, _cost_digest(self))
        cache = self.__dict__.setdefault("_probe", {})
        if key not in cache:
            for k in [k for k in cache if k[:2] == key[:2]]:
                cache.pop(k)
            cache[key] = BatchSolver(self.plant, self, None, N=2, dt=0.1, batch=1, qf_start_override=(0 if use_qf_running else -1))
        s = cache[key]
        s.set_goals(_as_f64(self.xg).reshape(1, -1))
        X = np.zeros((1, 2 * n, 2)); U = np.zeros((1, n, 1))
        xv = np.asarray(x, dtype=np.float64).reshape(-1)
        X[0, :, 0] = xv; X[0, :, 1] = xv
        if u is not None:
            U[0, :, 0] = np.asarray(u, dtype=np.float64).reshape(-1)
        s.set_trajectory(X, U)
        return s, (1 if u is None else 0)

    def _save(self, tag, **entries):
        for name, value in entries.items():
            getattr(self, "saved_" + name).append(dict(value=value, **tag))

    def _jtot(self, s, k):
        return s.fetch("cost_jtot")[0, k].reshape(2 * self.n, 2 * self.n)[:4].copy()

    def value(self, x, u=None, timestep=None, iter_1=0, iter_2=0, iter_3=0):
        s, k = self._knot(x, u, timestep)
        v = float(s.fetch("cost_value")[0, k, 0])
        if self.recording:
            self._save(dict(iteration=iter_1, outer_iteration=iter_2, line_search_iteration=iter_3), cost=v, dx=s.fetch("cost_err")[0, k, :4].copy())
        return v

    def gradient(self, x, u=None, timestep=None, iter_1=0, iter_2=0, iter_3=0):
        s, k = self._knot(x, u, timestep)
        g = s.fetch("cost_grad")[0, k]
        g = g[:2 * self.n].copy() if u is None else g.copy()
        if self.recording:
            self._save(dict(iteration=iter_1, outer_iteration=iter_2, line_search_iteration=iter_3), grad=g, Jacobian_tot_state=self._jtot(s, k))
        return g

    def hessian(self, x, u=None, timestep=None, iter_1=0, iter_2=0, iter_3=0):
        s, k = self._knot(x, u, timestep)
        m = 3 * self.n
        H = s.fetch("cost_hess")[0, k].reshape(m, m)
        H = H[:2 * self.n, :2 * self.n].copy() if u is None else H.copy()
        if self.recording:
            self._save(dict(iteration=iter_1, outer_iteration=iter_2, line_search_iteration=iter_3), hess=H, Jacobian_tot_state=self._jtot(s, k))
        return H

    def delta_x(self, x):
        """End-effector state error [ee_pos; J qd] - xg (TrajoptCost.py:425-435), evaluated by the cost kernel."""
        s, k = self._knot(x, None, None)
        return s.fetch("cost_err")[0, k, :4].copy()


# --------------------------------------------------------------------------------------------------- constraints
_MODES = {"QUADRATIC_PENALTY": _lib.LIMIT_QUADRATIC_PENALTY, "AUGMENTED_LAGRANGIAN": _lib.LIMIT_AUGMENTED_LAGRANGIAN,
          "ACTIVE_SET": _lib.LIMIT_ACTIVE_SET}


class BoxConstraint:
    """BoxConstraint (TrajoptConstraint.py:5-176) state holder: bounds, mode, options and the (mu, lambda, phi) arrays.
    The arithmetic lives in the kernels (csrc/b2t_core.cuh soft_value / soft_grad / hard_active, k_outer).  Soft modes
    QUADRATIC_PENALTY / AUGMENTED_LAGRANGIAN with every method; the hard mode ACTIVE_SET (violated bounds become KKT rows,
    TrajoptMPCReference.py:238-248) with the exact methods N / S."""

    def __init__(self, constraint_size=0, num_timesteps=0, upper_bounds=(), lower_bounds=(), mode="NONE", options=None):
        options = {} if options is None else options
        self.constraint_size = constraint_size
        self.num_timesteps = num_timesteps
        self.num_constraints = 2 * constraint_size * num_timesteps
        lblen, ublen = len(lower_bounds), len(upper_bounds)
        if (lblen != constraint_size and lblen != 1) or (ublen != constraint_size and ublen != 1):
            raise ValueError("[!]ERROR please enter bounds of the size of constraint or constant 1")
        self.bounds = np.zeros(2 * constraint_size)
        self.bounds[:constraint_size] = lower_bounds
        self.bounds[constraint_size:] = upper_bounds
        if mode == "FULL_SET":
            raise ValueError("FULL_SET adds the inactive bounds as all-zero KKT rows (TrajoptConstraint.py:66-67, 111-112): singular in the "
                             "reference itself; use ACTIVE_SET, QUADRATIC_PENALTY or AUGMENTED_LAGRANGIAN")
        if mode == "ADMM_PROJECTION":
            raise ValueError("[!] ERROR NOT IMPLEMENTED YET")            # same as the reference (TrajoptConstraint.py:87-89)
        if mode not in _MODES:
            raise ValueError("[!Error] Invalid Constraint Mode. Options are [ACTIVE_SET, FULL_SET, QUADRATIC_PENALTY, AUGMENTED_LAGRANGIAN, ADMM_PROJECTION]")
        self.mode = mode
        options.setdefault("quadratic_penalty_mu_init", 1e-2)
        options.setdefault("quadratic_penalty_mu_factor", 10.0)
        options.setdefault("quadratic_penalty_mu_max", 1e12)
        options.setdefault("augmentated_lagrangian_phi_init", 1e-2)
        options.setdefault("augmentated_lagrangian_phi_factor", 10.0)
        options.setdefault("jacobian_extra_columns_head", 0)
        options.setdefault("jacobian_extra_columns_tail", 0)
        self.options = options
        self.quadratic_penalty_mu = options["quadratic_penalty_mu_init"] * np.ones((2 * constraint_size, num_timesteps))
        self.augmented_lagrangian_lambda = np.zeros((2 * constraint_size, num_timesteps))
        self.augmented_lagrangian_phi = options["augmentated_lagrangian_phi_init"] * np.ones((2 * constraint_size, num_timesteps))

    def is_hard_constraint_mode(self, mode=None):
        return (mode or self.mode) in ["ACTIVE_SET", "FULL_SET"]

    def is_soft_constraint_mode(self, mode=None):
        return (mode or self.mode) in ["QUADRATIC_PENALTY", "AUGMENTED_LAGRANGIAN", "ADMM_PROJECTION"]

    def shift_soft_constraint_constants(self, shift_steps: int):
        """TrajoptConstraint.py:168-176 (literal, including which columns are re-initialised)."""
        self.quadratic_penalty_mu[:, :-shift_steps] = self.quadratic_penalty_mu[:, shift_steps:]
        self.augmented_lagrangian_lambda[:, :-shift_steps] = self.augmented_lagrangian_lambda[:, shift_steps:]
        self.augmented_lagrangian_phi[:, :-shift_steps] = self.augmented_lagrangian_phi[:, shift_steps:]
        self.quadratic_penalty_mu[:, shift_steps:] = self.options["quadratic_penalty_mu_init"]
        self.augmented_lagrangian_lambda[:, shift_steps:] = 0.0
        self.augmented_lagrangian_phi[:, shift_steps:] = self.options["augmentated_lagrangian_phi_init"]


class TrajoptConstraint:
    """TrajoptConstraint (TrajoptConstraint.py:178-387).  Joint and velocity limits carry N columns of multipliers, torque
    limits N-1 (the reference allocates N-1 for joint limits and then indexes column N-1)."""

    def __init__(self, nq: int = 0, nv: int = 0, nu: int = 0, num_timesteps: int = 0):
        self.nq, self.nv, self.nu, self.num_timesteps = nq, nv, nu, num_timesteps
        self.joint_limits = None
        self.velocity_limits = None
        self.torque_limits = None

    def set_joint_limits(self, upper_bounds, lower_bounds, mode, options=None):
        options = {} if options is None else options
        options["jacobian_extra_columns_tail"] = self.nv + self.nu
        self.joint_limits = BoxConstraint(self.nq, self.num_timesteps, upper_bounds, lower_bounds, mode, options)

    def set_velocity_limits(self, upper_bounds, lower_bounds, mode, options=None):
        options = {} if options is None else options
        options["jacobian_extra_columns_head"] = self.nq
        options["jacobian_extra_columns_tail"] = self.nu
        self.velocity_limits = BoxConstraint(self.nv, self.num_timesteps, upper_bounds, lower_bounds, mode, options)

    def set_torque_limits(self, upper_bounds, lower_bounds, mode, options=None):
        options = {} if options is None else options
        options["jacobian_extra_columns_head"] = self.nq + self.nv
        self.torque_limits = BoxConstraint(self.nu, self.num_timesteps - 1, upper_bounds, lower_bounds, mode, options)

    def _types(self):
        return [(0, self.joint_limits, 0, self.nq), (1, self.velocity_limits, self.nq, self.nv),
                (2, self.torque_limits, self.nq + self.nv, self.nu)]

    def total_soft_constraints(self, timestep=None):
        total = 0
        for ty, lim, off, cs in self._types():
            if lim is None:
                continue
            if timestep is None:
                total += lim.num_constraints
            elif not (ty == 2 and timestep >= self.num_timesteps - 1):
                total += lim.constraint_size
        return total

    def shift_soft_constraint_constants(self, shift_steps: int):
        for _, lim, _, _ in self._types():
            if lim is not None:
                lim.shift_soft_constraint_constants(shift_steps)

    # ---- per-knot callbacks (TrajoptConstraint.py:295-340), evaluated by the constraint kernels with this object's multipliers
    def bind_plant(self, plant):
        """The callbacks run on the robot's CUDA library: tell the constraint object which plant it belongs to (TrajoptMPCReference
        does this for the objects it is given)."""
        self._plant = plant
        return self

    def _knot(self, xk, uk, timestep):
        plant = getattr(self, "_plant", None)
        if plant is None:
            raise ValueError("call bind_plant(plant) first (or pass the object to TrajoptMPCReference): the constraint callbacks run on the GPU")
        N = self.num_timesteps
        if timestep is None:
            timestep = N - 1
        cache = self.__dict__.setdefault("_probe", {})
        dig = (_constraint_digest(self), id(plant))
        if cache.get("key") != dig:
            n = self.nq
            cache["s"] = BatchSolver(plant, QuadraticCost(np.eye(2 * n), np.eye(2 * n), np.eye(n), np.zeros(2 * n)), self, N=N, dt=0.1, batch=1)
            cache["key"] = dig
        s = cache["s"]
        mu, lam, phi = self.pack(N)
        s.set_multipliers(mu[None], lam[None], phi[None])
        X = np.zeros((1, 2 * self.nq, N)); U = np.zeros((1, self.nu, N - 1))
        X[0, :, timestep] = np.asarray(xk, dtype=np.float64).reshape(-1)
        if uk is not None and timestep < N - 1:
            U[0, :, timestep] = np.asarray(uk, dtype=np.float64).reshape(-1)
        s.set_trajectory(X, U)
        return s, timestep

    def value_soft_constraints(self, xk, uk=None, timestep=None):
        s, k = self._knot(xk, uk, timestep)
        return float(s.fetch("soft_value")[0, k, 0])

    def jacobian_soft_constraints(self, xk, uk=None, timestep=None):
        """Summed penalty gradient gck as an (m, 1) column (what SQP adds to g_k, TrajoptMPCReference.py:220-224)."""
        s, k = self._knot(xk, uk, timestep)
        return s.fetch("soft_grad")[0, k].reshape(-1, 1).copy()

    # ---- device layout helpers: [2m][N] per instance, lower coordinate i -> row i, upper -> row m + i
    def pack(self, N):
        m = self.nq + self.nv + self.nu
        mu = np.zeros((2 * m, N)); lam = np.zeros((2 * m, N)); phi = np.ones((2 * m, N))
        for ty, lim, off, cs in self._types():
            if lim is None:
                continue
            T = lim.num_timesteps
            for arr, src in ((mu, lim.quadratic_penalty_mu), (lam, lim.augmented_lagrangian_lambda), (phi, lim.augmented_lagrangian_phi)):
                arr[off:off + cs, :T] = src[:cs]
                arr[m + off:m + off + cs, :T] = src[cs:]
        return mu, lam, phi

    def unpack(self, mu, lam, phi):
        m = self.nq + self.nv + self.nu
        for ty, lim, off, cs in self._types():
            if lim is None or not lim.is_soft_constraint_mode():
                continue
            T = lim.num_timesteps
            for arr, dst in ((mu, lim.quadratic_penalty_mu), (lam, lim.augmented_lagrangian_lambda), (phi, lim.augmented_lagrangian_phi)):
                dst[:cs] = arr[off:off + cs, :T]
                dst[cs:] = arr[m + off:m + off + cs, :T]


# --------------------------------------------------------------------------------------------------- batched solver
def _dptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _as_f64(a, shape=None):
    a = np.ascontiguousarray(np.asarray(a, dtype=np.float64))
    if shape is not None and a.shape != tuple(shape):
        raise ValueError("expected array of shape %r, got %r" % (tuple(shape), a.shape))
    return a


def _cost_digest(cost):
    """Content digest of everything a BatchSolver uploads ONCE from a cost object (Q, QF, R, QF_start, kind).  The reference's
    in-place mutators (increase_QF / increase_Q / shift_QF_start, TrajoptCost.py:85-104) change these after a solver was cached;
    xg is re-sent on every call and is not part of the digest."""
    h = hashlib.sha1()
    for a in (cost.Q, cost.QF, cost.R):
        h.update(np.ascontiguousarray(np.asarray(a, dtype=np.float64)).tobytes())
    h.update(repr((cost._kind, cost.QF_start, getattr(cost, "hess_mode", 0))).encode())
    return h.hexdigest()


def _constraint_digest(cons):
    """Content digest of the limit description a BatchSolver uploads once: per limit type the bounds, mode and the mu / phi
    schedule options (not the multipliers: those are re-sent per call).  set_*_limits after a first solve changes it."""
    if cons is None:
        return "none"
    h = hashlib.sha1()
    h.update(repr((cons.nq, cons.nv, cons.nu, cons.num_timesteps)).encode())
    for ty, lim, off, cs in cons._types():
        if lim is None:
            h.update(b"-")
            continue
        h.update(np.ascontiguousarray(np.asarray(lim.bounds, dtype=np.float64)).tobytes())
        o = lim.options
        h.update(repr((ty, lim.mode, lim.num_timesteps, o["quadratic_penalty_mu_init"], o["quadratic_penalty_mu_factor"], o["quadratic_penalty_mu_max"],
                       o["augmentated_lagrangian_phi_init"], o["augmentated_lagrangian_phi_factor"])).encode())
    return h.hexdigest()


class BatchResult(dict):
    __getattr__ = dict.__getitem__


class BatchSolver:
    """One device workspace for `batch` independent instances of (plant, cost, constraints, N, dt)."""

    def __init__(self, plant, cost, constraints, N, dt, batch=1, dtype="f64", device=0, qf_start_override=None, dense_kkt=False,
                 integrator_type=None):
        if not isinstance(plant, URDFPlant):
            raise ValueError("Must pass in a URDFPlant: the dynamics kernels are generated from the URDF")
        if not isinstance(cost, QuadraticCost):
            raise ValueError("cost must be a QuadraticCost or UrdfCost (Python callbacks cannot run inside the kernels)")
        n = plant.get_num_pos()
        self.plant, self.cost, self.constraints = plant, cost, constraints
        self.n, self.nx, self.nu, self.m = n, 2 * n, n, 3 * n
        self.N, self.dt, self.batch, self.dtype, self.device = int(N), float(dt), int(batch), dtype, int(device)
        self.lib = plant.lib
        d = _lib.ProblemDesc()
        self.integrator_type = plant.integrator_type if integrator_type is None else int(integrator_type)
        if self.integrator_type == 4:
            raise TypeError("integrator type 4 (rk4): the reference's gradient branch raises TypeError (TrajoptPlant.py:259), so no solve can run")
        d.batch, d.knots, d.integrator_type = self.batch, self.N, self.integrator_type
        d.dtype = {"f64": _lib.F64, "f32": _lib.F32}[dtype]
        d.dt, d.gravity = self.dt, float(plant.options["gravity"])
        d.cost_kind = cost._kind
        qs = cost.QF_start if qf_start_override is None else qf_start_override
        d.qf_start = -1 if qs is None else int(qs)
        d.hess_mode = int(getattr(cost, "hess_mode", 0))
        self.ne = 4 if cost._kind == _lib.COST_URDF_EE else self.nx      # size of the cost's error vector / of Q, QF, xg

        def padded(M):      # the C ABI takes nx*nx doubles; an end-effector cost packs its 4 x 4 weights in the first 16
            out = np.zeros(self.nx * self.nx)
            out[:self.ne * self.ne] = _as_f64(M, (self.ne, self.ne)).reshape(-1)
            return out
        self._keep = [padded(cost.Q), padded(cost.QF), _as_f64(cost.R, (self.nu, self.nu))]
        d.Q, d.QF, d.R = [a.ctypes.data_as(ctypes.POINTER(ctypes.c_double)) for a in self._keep]
        lower = np.zeros(self.m); upper = np.zeros(self.m)
        self.has_limits = False
        self.has_soft_limits = False          # penalty / augmented-Lagrangian limits carry (mu, lambda, phi) on the device
        for ty in range(3):
            d.limit_mode[ty] = _lib.LIMIT_NONE
            d.mu_init[ty], d.mu_factor[ty], d.mu_max[ty], d.phi_init[ty], d.phi_factor[ty] = 1e-2, 10.0, 1e12, 1e-2, 10.0
        if constraints is not None:
            if constraints.num_timesteps != self.N:
                raise ValueError("TrajoptConstraint.num_timesteps must equal N")
            for ty, lim, off, cs in constraints._types():
                if lim is None:
                    continue
                self.has_limits = True
                self.has_soft_limits = self.has_soft_limits or lim.is_soft_constraint_mode()
                d.limit_mode[ty] = _MODES[lim.mode]
                lower[off:off + cs] = lim.bounds[:cs]; upper[off:off + cs] = lim.bounds[cs:]
                o = lim.options
                d.mu_init[ty], d.mu_factor[ty], d.mu_max[ty] = o["quadratic_penalty_mu_init"], o["quadratic_penalty_mu_factor"], o["quadratic_penalty_mu_max"]
                d.phi_init[ty], d.phi_factor[ty] = o["augmentated_lagrangian_phi_init"], o["augmentated_lagrangian_phi_factor"]
        self._keep += [lower, upper]
        d.lower = lower.ctypes.data_as(ctypes.POINTER(ctypes.c_double)); d.upper = upper.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
        self._h = ctypes.c_void_p()
        # dense_kkt=True forces the general (dense G_k) kernels even when Q, QF, R are diagonal (used by the parity tests)
        old = os.environ.get("B2T_DENSE_KKT")
        if dense_kkt:
            os.environ["B2T_DENSE_KKT"] = "1"
        try:
            _lib.check(self.lib, self.lib.b2t_solver_create(ctypes.byref(d), self.device, ctypes.byref(self._h)))
        finally:
            if dense_kkt:
                if old is None:
                    os.environ.pop("B2T_DENSE_KKT", None)
                else:
                    os.environ["B2T_DENSE_KKT"] = old
        xg = np.broadcast_to(_as_f64(cost.xg).reshape(1, -1), (self.batch, self.ne))
        self.set_goals(xg)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self.lib.b2t_solver_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def workspace_bytes(self):
        return int(self.lib.b2t_workspace_bytes(self._h))

    # ---- inputs
    def set_trajectory(self, x, u):
        """x (batch, nx, N), u (batch, nu, N-1): numpy (host) or torch CUDA float64 tensors (device, zero-copy)."""
        if hasattr(x, "data_ptr"):
            assert x.is_cuda and u.is_cuda and x.is_contiguous() and u.is_contiguous() and str(x.dtype) == "torch.float64"
            assert tuple(x.shape) == (self.batch, self.nx, self.N) and tuple(u.shape) == (self.batch, self.nu, self.N - 1)
            _lib.check(self.lib, self.lib.b2t_set_trajectory(self._h, ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(u.data_ptr()), 1))
            return
        x = _as_f64(x, (self.batch, self.nx, self.N)); u = _as_f64(u, (self.batch, self.nu, self.N - 1))
        _lib.check(self.lib, self.lib.b2t_set_trajectory(self._h, _dptr(x), _dptr(u), 0))
        self._sync_keep = (x, u)

    def set_goals(self, xg):
        if hasattr(xg, "data_ptr"):
            assert xg.is_cuda and xg.is_contiguous() and tuple(xg.shape) == (self.batch, self.nx)
            _lib.check(self.lib, self.lib.b2t_set_goals(self._h, ctypes.c_void_p(xg.data_ptr()), 1))
            return
        xg = _as_f64(xg, (self.batch, self.ne))
        if self.ne != self.nx:      # end-effector goals (x, y, vx, vy) occupy the first 4 of the nx slots per instance
            xg = np.concatenate([xg, np.zeros((self.batch, self.nx - self.ne))], axis=1)
            xg = np.ascontiguousarray(xg)
        _lib.check(self.lib, self.lib.b2t_set_goals(self._h, _dptr(xg), 0))
        self._goal_keep = xg

    def set_initial_state(self, xs):
        xs = _as_f64(xs, (self.batch, self.nx))
        _lib.check(self.lib, self.lib.b2t_set_initial_state(self._h, _dptr(xs)))

    def set_multipliers(self, mu, lam, phi):
        if not self.has_soft_limits:          # hard (ACTIVE_SET) limits only: nothing to send
            return
        args = [_as_f64(a, (self.batch, 2 * self.m, self.N)) for a in (mu, lam, phi)]
        _lib.check(self.lib, self.lib.b2t_set_multipliers(self._h, *[_dptr(a) for a in args]))

    def get_multipliers(self):
        if not self.has_soft_limits:
            shape = (self.batch, 2 * self.m, self.N)
            return np.zeros(shape), np.zeros(shape), np.ones(shape)
        out = [np.zeros((self.batch, 2 * self.m, self.N)) for _ in range(3)]
        _lib.check(self.lib, self.lib.b2t_get_multipliers(self._h, *[_dptr(a) for a in out]))
        return out

    def reset_multipliers(self):
        _lib.check(self.lib, self.lib.b2t_reset_multipliers(self._h))

    # ---- solve
    def make_options(self, options=None):
        o = _lib.Options()
        self.lib.b2t_default_options(ctypes.byref(o))
        options = options or {}
        mapping = {"exit_tolerance_linSys": "exit_tolerance_linSys", "max_iter_linSys": "max_iter_linSys",
                   "exit_tolerance_SQP_DDP": "exit_tolerance_SQP", "max_iter_SQP_DDP": "max_iter_SQP",
                   "alpha_factor_SQP_DDP": "alpha_factor", "alpha_min_SQP_DDP": "alpha_min", "rho_factor_SQP_DDP": "rho_factor",
                   "rho_min_SQP_DDP": "rho_min", "rho_max_SQP_DDP": "rho_max", "rho_init_SQP_DDP": "rho_init",
                   "expected_reduction_min_SQP_DDP": "expected_reduction_min", "expected_reduction_max_SQP_DDP": "expected_reduction_max",
                   "exit_tolerance_softConstraints": "exit_tolerance_soft", "max_iter_softConstraints": "max_iter_soft"}
        for k, f in mapping.items():
            if k in options:
                setattr(o, f, type(getattr(o, f))(options[k]))
        return o

    def solve(self, method=SQPSolverMethods.PCG_SS, options=None):
        """Runs SQP on the trajectories / goals currently in the workspace."""
        if method not in _METHOD_CODE:
            raise ValueError("Invalid QP Solver options are: N, S, PCG-J, PCG-BJ, PCG-SS")
        o = self.make_options(options)
        code = self.lib.b2t_sqp_solve(self._h, _METHOD_CODE[method], ctypes.byref(o))
        if getattr(self, "_hook_error", None) is not None:
            exc, self._hook_error = self._hook_error, None
            raise exc
        _lib.check(self.lib, code)

    def set_iteration_hook(self, fn):
        """fn(event, pass) -> None, called on the host inside solve() after the linear solve (event _lib.HOOK_LINSYS) and after the step
        (event _lib.HOOK_STEP) of every SQP iteration; `fetch`, `get_status`, `get_trajectory` may be called from it.  None removes it."""
        if fn is None:
            self._hook = _lib.ITERATION_HOOK(0)
        else:
            def tramp(_user, event, ipass):
                try:
                    fn(int(event), int(ipass))
                    return 0
                except BaseException as exc:       # cannot propagate through C: stop the solve and re-raise afterwards
                    self._hook_error = exc
                    return 1
            self._hook = _lib.ITERATION_HOOK(tramp)
        self._hook_error = None
        _lib.check(self.lib, self.lib.b2t_set_iteration_hook(self._h, self._hook, None))

    def solve_ilqr(self, options=None):
        """iLQR on the trajectories / goals currently in the workspace (x[:,0] = start state, x re-rolled from u)."""
        o = self.make_options(options)
        _lib.check(self.lib, self.lib.b2t_ilqr_solve(self._h, ctypes.byref(o)))

    def mpc_shift(self, x_next=None):
        """Receding-horizon shift by one knot; returns (x_0, u_0 applied, next initial state).  x_next=None simulates the plant."""
        x0 = np.zeros((self.batch, self.nx)); u0 = np.zeros((self.batch, self.nu)); xn = np.zeros((self.batch, self.nx))
        arg = None if x_next is None else _dptr(_as_f64(x_next, (self.batch, self.nx)))
        _lib.check(self.lib, self.lib.b2t_mpc_shift(self._h, arg, _dptr(x0), _dptr(u0), _dptr(xn)))
        return x0, u0, xn

    def solve_host(self, x0, u0, xg, x_out, u_out, status_out, method=SQPSolverMethods.PCG_SS, options=None):
        """One call: host buffers in, host buffers out (pinned buffers make the copies asynchronous)."""
        o = self.make_options(options)
        _lib.check(self.lib, self.lib.b2t_sqp_solve_host(self._h, _dptr(x0), _dptr(u0), None if xg is None else _dptr(xg),
                                                        _METHOD_CODE[method], ctypes.byref(o), _dptr(x_out), _dptr(u_out), _dptr(status_out)))

    def get_trajectory(self, x_out=None, u_out=None):
        if x_out is not None and hasattr(x_out, "data_ptr"):
            _lib.check(self.lib, self.lib.b2t_get_trajectory(self._h, ctypes.c_void_p(x_out.data_ptr()), ctypes.c_void_p(u_out.data_ptr()), 1))
            return x_out, u_out
        x = np.zeros((self.batch, self.nx, self.N)) if x_out is None else x_out
        u = np.zeros((self.batch, self.nu, self.N - 1)) if u_out is None else u_out
        _lib.check(self.lib, self.lib.b2t_get_trajectory(self._h, _dptr(x), _dptr(u), 0))
        return x, u

    def get_status(self):
        st = np.zeros((self.batch, _lib.STATUS_FIELDS), dtype=np.int32)
        _lib.check(self.lib, self.lib.b2t_get_status(self._h, _dptr(st)))
        return st

    def get_scalars(self):
        sc = np.zeros((self.batch, _lib.SCALAR_FIELDS))
        _lib.check(self.lib, self.lib.b2t_get_scalars(self._h, _dptr(sc)))
        return sc

    def get_trace(self, cap=104):
        tr = np.zeros((self.batch, cap, _lib.TRACE_FIELDS))
        _lib.check(self.lib, self.lib.b2t_get_trace(self._h, _dptr(tr), cap))
        return tr

    def launch_stats(self):
        n = ctypes.c_longlong(); s = ctypes.c_double()
        _lib.check(self.lib, self.lib.b2t_get_launch_stats(self._h, ctypes.byref(n), ctypes.byref(s)))
        return int(n.value), float(s.value)

    def pass_trace(self):
        """(number of SQP passes of the last solve, active-instance count after each pass)."""
        n = ctypes.c_int()
        _lib.check(self.lib, self.lib.b2t_get_pass_trace(self._h, None, 0, ctypes.byref(n)))
        buf = (ctypes.c_int * max(1, n.value))()
        _lib.check(self.lib, self.lib.b2t_get_pass_trace(self._h, buf, n.value, ctypes.byref(n)))
        return n.value, np.array(buf[:min(n.value, 2048)], dtype=np.int64)

    def pcg_kernel_name(self):
        return self.lib.b2t_pcg_kernel_name(self._h).decode()

    def set_profiling(self, enabled, family=None):
        """CUDA-event timing of the kernel families of the next solves: all of them, or only `family` (a name of
        _lib.KERNEL_FAMILY_NAMES; two event records per launch cost ~2 % of a step when every family is timed)."""
        mode = 0 if not enabled else (1 if family is None else 2 + _lib.KERNEL_FAMILY_NAMES.index(family))
        _lib.check(self.lib, self.lib.b2t_set_profiling(self._h, mode))

    def kernel_times(self):
        sec = (ctypes.c_double * _lib.KERNEL_FAMILIES)(); cnt = (ctypes.c_longlong * _lib.KERNEL_FAMILIES)()
        _lib.check(self.lib, self.lib.b2t_get_kernel_times(self._h, sec, cnt))
        return {name: (float(sec[i]), int(cnt[i])) for i, name in enumerate(_lib.KERNEL_FAMILY_NAMES)}

    def result(self):
        x, u = self.get_trajectory()
        st = self.get_status(); sc = self.get_scalars()
        return BatchResult(x=x, u=u, exit_sqp=st[:, 0], exit_soft=st[:, 1], outer_iter=st[:, 2], sqp_iter=st[:, 3], total_qp=st[:, 4],
                           total_pcg=st[:, 5], total_trials=st[:, 6], trace_rows=st[:, 7], J=sc[:, 0], c=sc[:, 1], merit=sc[:, 2], rho=sc[:, 3])

    # ---- stages (parity tests)
    def stage_dynamics(self):
        _lib.check(self.lib, self.lib.b2t_stage_dynamics(self._h))

    def stage_kkt(self, rho, method=SQPSolverMethods.PCG_SS):
        _lib.check(self.lib, self.lib.b2t_stage_kkt(self._h, float(rho), _METHOD_CODE[method]))

    def stage_pcg(self, method=SQPSolverMethods.PCG_SS, tol=1e-6, max_iter=100):
        it = np.zeros(self.batch, dtype=np.int32)
        _lib.check(self.lib, self.lib.b2t_stage_pcg(self._h, _METHOD_CODE[method], float(tol), int(max_iter), _dptr(it)))
        return it

    def set_block_system(self, Sd, So, gamma):
        """Upload a block-tridiagonal system (batch, N, nx*nx) x2, (batch, N, nx) for the standalone PCG entry."""
        a = _as_f64(Sd, (self.batch, self.N, self.nx * self.nx)); b = _as_f64(So, (self.batch, self.N, self.nx * self.nx))
        c = _as_f64(gamma, (self.batch, self.N, self.nx))
        _lib.check(self.lib, self.lib.b2t_set_block_system(self._h, _dptr(a), _dptr(b), _dptr(c)))

    def stage_precond(self, method=SQPSolverMethods.PCG_SS):
        _lib.check(self.lib, self.lib.b2t_stage_precond(self._h, _METHOD_CODE[method]))

    def stage_recover(self):
        _lib.check(self.lib, self.lib.b2t_stage_recover(self._h))

    def stage_merit(self, alpha):
        J = np.zeros(self.batch); c = np.zeros(self.batch); D = np.zeros(self.batch)
        _lib.check(self.lib, self.lib.b2t_stage_merit(self._h, float(alpha), _dptr(J), _dptr(c), _dptr(D)))
        return J, c, D

    def measure_fma_peak(self, dtype="f64"):
        """Measured FMA throughput (TFLOP/s) of this GPU's fp64 / fp32 pipe: the roofline denominator of the path."""
        v = ctypes.c_double()
        _lib.check(self.lib, self.lib.b2t_measure_fma_peak(self.device, {"f64": _lib.F64, "f32": _lib.F32}[dtype], ctypes.byref(v)))
        return float(v.value)

    def fetch(self, name):
        """Knot-major copy of an internal array: (batch, N, elems)."""
        E = {"x": self.nx, "u": self.nu, "xkp1": self.nx, "dqdd": self.n * 3 * self.n, "Ghat": self.m * self.m, "g": self.m,
             "Sd": self.nx * self.nx, "So": self.nx * self.nx, "Pd": self.nx * self.nx, "gamma": self.nx, "l": self.nx, "dz": self.m,
             "xn": self.nx, "un": self.nu, "cost_value": 1, "cost_grad": self.m, "cost_hess": self.m * self.m, "cost_err": self.nx,
             "kkt_hess": self.m * self.m, "AB": self.nx * self.m, "soft_value": 1, "soft_grad": self.m, "nu_trace": None,
             "cost_jtot": self.nx * self.nx, "plant_terms": 2 * self.n + 3 * self.n * self.n}[name]
        if name == "nu_trace":
            out = np.zeros((self.batch, 128))
            _lib.check(self.lib, self.lib.b2t_fetch(self._h, _lib.ARR[name], _dptr(out)))
            return out
        out = np.zeros((self.batch, self.N, E))
        _lib.check(self.lib, self.lib.b2t_fetch(self._h, _lib.ARR[name], _dptr(out)))
        return out


_TRACE_KEYS = ["outer_iteration", "iteration", "line_search_iteration", "alpha", "rho", "J", "c", "merit", "D", "reduction_ratio",
               "inner_iters", "succeeded_line_search"]


class TrajoptMPCReference:
    """TrajoptMPCReference (TrajoptMPCReference.py:29-760), SQP with the Schur-complement / GBD-PCG linear solve."""

    def __init__(self, plantObj, costObj, constraintObj=None):
        if not isinstance(plantObj, TrajoptPlant) or not isinstance(costObj, TrajoptCost):
            raise ValueError("Must pass in a TrajoptPlant and TrajoptCost object to TrajoptMPCReference.")
        if constraintObj is None:
            constraintObj = TrajoptConstraint()
        elif not isinstance(constraintObj, TrajoptConstraint):
            raise ValueError("If passing in additional constraints must pass in a TrajoptConstraint object to TrajoptMPCReference.")
        self.plant, self.cost, self.other_constraints = plantObj, costObj, constraintObj
        if isinstance(plantObj, URDFPlant):
            constraintObj.bind_plant(plantObj)
        self.trace = []
        self.exit_soft = 0
        self.exit_sqp = 0
        self.singular = False
        self.n_inner_iter = 0
        self.pcg_iters = []
        self._solvers = {}

    def update_cost(self, costObj):
        assert isinstance(costObj, TrajoptCost), "Must pass in a TrajoptCost object to update_cost in TrajoptMPCReference."
        self.cost = costObj
        self._solvers.clear()

    def update_plant(self, plantObj):
        assert isinstance(plantObj, TrajoptPlant), "Must pass in a TrajoptPlant object to update_plant in TrajoptMPCReference."
        self.plant = plantObj
        self._solvers.clear()

    def update_constraints(self, constraintObj):
        assert isinstance(constraintObj, TrajoptConstraint), "Must pass in a TrajoptConstraint object to update_constraints in TrajoptMPCReference."
        self.other_constraints = constraintObj
        if isinstance(self.plant, URDFPlant):
            constraintObj.bind_plant(self.plant)
        self._solvers.clear()

    def set_default_options(self, options: dict):
        """TrajoptMPCReference.set_default_options (:91-115): fills the caller's dict in place."""
        options.setdefault("exit_tolerance_linSys", 1e-6)
        options.setdefault("max_iter_linSys", 100)
        options.setdefault("DEBUG_MODE_linSys", False)
        options.setdefault("RETURN_TRACE_linSys", False)
        options.setdefault("overloading", self.plant.rbdReference.overloading)
        options.setdefault("exit_tolerance_SQP_DDP", 1e-6)
        options.setdefault("max_iter_SQP_DDP", 100)
        options.setdefault("DEBUG_MODE_SQP_DDP", False)
        options.setdefault("alpha_factor_SQP_DDP", 0.5)
        options.setdefault("alpha_min_SQP_DDP", 0.005)
        options.setdefault("rho_factor_SQP_DDP", 4)
        options.setdefault("rho_min_SQP_DDP", 1e-3)
        options.setdefault("rho_max_SQP_DDP", 1e3)
        options.setdefault("rho_init_SQP_DDP", 0.001)
        options.setdefault("expected_reduction_min_SQP_DDP", 0.05)
        options.setdefault("expected_reduction_max_SQP_DDP", 3)
        options.setdefault("merit_factor_SQP", 1.5)
        options.setdefault("exit_tolerance_softConstraints", 1e-6)
        options.setdefault("max_iter_softConstraints", 10)
        options.setdefault("DEBUG_MODE_Soft_Constraints", False)

    def _constraints_or_none(self):
        c = self.other_constraints
        if c is None or all(l is None for l in (c.joint_limits, c.velocity_limits, c.torque_limits)):
            return None
        return c

    def batch_solver(self, N, dt, batch, dtype="f64", device=0):
        # keyed on the CONTENT the workspace uploads once (weights, QF_start, bounds, modes, penalty options), not on object identity:
        # the reference API mutates costs and limits in place (increase_QF, shift_QF_start, set_*_limits after a first solve)
        cons = self._constraints_or_none()
        key = (N, float(dt), batch, dtype, device, self.plant.integrator_type, _cost_digest(self.cost), _constraint_digest(cons))
        if key not in self._solvers:
            shape = key[:5]
            for k in [k for k in self._solvers if k[:5] == shape]:      # same shape, stale content: drop it (freed when unreferenced)
                self._solvers.pop(k)
            self._solvers[key] = BatchSolver(self.plant, self.cost, cons, N, dt, batch, dtype, device)
        return self._solvers[key]

    def SQP(self, x, u, N, dt, LINEAR_SYSTEM_SOLVER_METHOD=SQPSolverMethods.N, options=None, dtype="f64", record=False):
        """Same call and return value as the reference (:510, :760): (x, u, exit_sqp, exit_soft, outer_iter, sqp_iter).
        record=True additionally fills the reference's `saved_*` lists (dense G, g, C, c, invG, S, gamma, Pinv, l, dxul, A_k, B_k, x, u
        per SQP iteration) from the device through the iteration hook; see record.py."""
        options = {} if options is None else options
        self.set_default_options(options)
        if not isinstance(LINEAR_SYSTEM_SOLVER_METHOD, SQPSolverMethods):
            raise ValueError("Invalid QP Solver options are: N, S, PCG-J, PCG-BJ, PCG-SS")
        if options.get("overloading"):
            raise ValueError("the operator-overloading tracer (overloading.py) is research instrumentation and is not supported")
        s = self.batch_solver(N, dt, 1, dtype)
        cons = self._constraints_or_none()
        x = _as_f64(x, (s.nx, N)); u = _as_f64(u, (s.nu, N - 1))
        s.set_goals(_as_f64(self.cost.xg).reshape(1, -1))
        s.set_trajectory(x[None], u[None])
        if cons is not None:
            mu, lam, phi = cons.pack(N)
            s.set_multipliers(mu[None], lam[None], phi[None])
        if record:
            from .record import Recorder
            s.set_iteration_hook(Recorder(self, s, LINEAR_SYSTEM_SOLVER_METHOD, x[:, 0], options))
        try:
            s.solve(LINEAR_SYSTEM_SOLVER_METHOD, options)
        finally:
            if record:
                s.set_iteration_hook(None)
        r = s.result()
        if cons is not None:
            mu, lam, phi = s.get_multipliers()
            cons.unpack(mu[0], lam[0], phi[0])
        self.exit_sqp, self.exit_soft = int(r.exit_sqp[0]), int(r.exit_soft[0])
        rows = int(r.trace_rows[0])
        tr = s.get_trace()[0][:min(rows, 104)]
        self.trace = []
        self.pcg_iters = []
        for i, row in enumerate(tr):
            dct = {k: row[j] for j, k in enumerate(_TRACE_KEYS)}
            for k in ("outer_iteration", "iteration", "line_search_iteration"):
                dct[k] = int(dct[k])
            dct["pcg_iters"] = int(dct["inner_iters"])
            dct["inner_iters"] = 0 if (i == 0 and dct["outer_iteration"] == 0) else 2      # reference quirk (:445): len of a 2-tuple
            dct["succeeded_line_search"] = bool(dct["succeeded_line_search"])
            dct["singular"] = False
            if i == 0:
                dct["D"] = None
                dct["reduction_ratio"] = None
                dct["alpha"] = 1
            else:
                self.pcg_iters.append(dct["pcg_iters"])
            self.trace.append(dct)
        self.n_inner_iter = 2
        self.last_result = r
        return r.x[0], r.u[0], self.exit_sqp, self.exit_soft, int(r.outer_iter[0]), int(r.sqp_iter[0])

    def iLQR(self, x, u, N, dt, options=None, dtype="f64"):
        """iLQR with soft (penalty / augmented-Lagrangian) box limits: MPCSolverMethods.iLQR, which the reference names
        (README.md:15-17) but does not implement.  Same options and return tuple as SQP; specification: oracle/ilqr.py."""
        options = {} if options is None else options
        self.set_default_options(options)
        r = self.ilqr_batch(_as_f64(x)[None], _as_f64(u)[None], _as_f64(self.cost.xg).reshape(1, -1), N, dt, options, dtype, single=True)
        self.exit_sqp, self.exit_soft = int(r.exit_sqp[0]), int(r.exit_soft[0])
        self.last_result = r
        return r.x[0], r.u[0], self.exit_sqp, self.exit_soft, int(r.outer_iter[0]), int(r.sqp_iter[0])

    def ilqr_batch(self, x0, u0, xg, N, dt, options=None, dtype="f64", single=False, device=0):
        options = {} if options is None else options
        self.set_default_options(options)
        B = x0.shape[0]
        s = self.batch_solver(N, dt, B, dtype, device)
        s.set_goals(xg)
        s.set_trajectory(x0, u0)
        cons = self._constraints_or_none()
        if cons is not None:
            mu, lam, phi = cons.pack(N)
            s.set_multipliers(*[np.broadcast_to(a[None], (B,) + a.shape) for a in (mu, lam, phi)])
        s.solve_ilqr(options)
        r = s.result()
        if single and cons is not None:
            mu, lam, phi = s.get_multipliers()
            cons.unpack(mu[0], lam[0], phi[0])
        return r

    def mpc_batch(self, x_start, xg, N, dt, steps, method=SQPSolverMethods.PCG_SS, options=None, use_ilqr=False, dtype="f64", device=0,
                  measure=None):
        """Receding-horizon MPC (README.md:15-17 names it; no reference code -- semantics in DESIGN.md section 9): `steps` times
        {solve from the warm start, apply u_0, obtain the next state (simulated by the plant, or measure(step, x, u) -> (B,nx)),
        shift trajectories and multipliers one knot}.  Everything stays on the device between solves.
        x_start (B, nx), xg (B, nx).  Returns BatchResult(x_closed (B,nx,steps+1), u_applied (B,nu,steps), sqp_iter (B,steps), J (B,steps))."""
        options = {} if options is None else options
        self.set_default_options(options)
        x_start = _as_f64(x_start); B = x_start.shape[0]
        s = self.batch_solver(N, dt, B, dtype, device)
        s.set_goals(xg)
        X0 = np.repeat(x_start[:, :, None], N, axis=2); U0 = np.zeros((B, s.nu, N - 1))
        s.set_trajectory(X0, U0)
        cons = self._constraints_or_none()
        if cons is not None:
            mu, lam, phi = cons.pack(N)
            s.set_multipliers(*[np.broadcast_to(a[None], (B,) + a.shape) for a in (mu, lam, phi)])
        xc = np.zeros((B, s.nx, steps + 1)); ua = np.zeros((B, s.nu, steps)); its = np.zeros((B, steps), dtype=np.int64); Js = np.zeros((B, steps))
        xc[:, :, 0] = x_start
        for k in range(steps):
            if use_ilqr:
                s.solve_ilqr(options)
            else:
                s.solve(method, options)
            st = s.get_status(); sc = s.get_scalars()
            its[:, k] = st[:, 3]; Js[:, k] = sc[:, 0]
            if measure is None:
                x0, u0, xn = s.mpc_shift(None)
            else:
                xt, ut = s.get_trajectory()
                xn = _as_f64(measure(k, xt[:, :, 0], ut[:, :, 0]), (B, s.nx))
                x0, u0, _ = s.mpc_shift(xn)
            ua[:, :, k] = u0; xc[:, :, k + 1] = xn
        return BatchResult(x_closed=xc, u_applied=ua, sqp_iter=its, J=Js)

    def solve_batch(self, x0, u0, xg, N, dt, LINEAR_SYSTEM_SOLVER_METHOD=SQPSolverMethods.PCG_SS, options=None, dtype="f64", device=0):
        """Batched SQP: x0 (B, nx, N), u0 (B, nu, N-1), xg (B, nx).  Returns a BatchResult of per-instance arrays."""
        options = {} if options is None else options
        self.set_default_options(options)
        B = x0.shape[0]
        s = self.batch_solver(N, dt, B, dtype, device)
        s.set_goals(xg)
        s.set_trajectory(x0, u0)
        cons = self._constraints_or_none()
        if cons is not None:
            mu, lam, phi = cons.pack(N)
            s.set_multipliers(*[np.broadcast_to(a[None], (B,) + a.shape) for a in (mu, lam, phi)])
        s.solve(LINEAR_SYSTEM_SOLVER_METHOD, options)
        return s.result()

    def totalCost(self, x, u, N):
        """totalCost (:296-310) of one trajectory, evaluated by the merit kernel."""
        s = self.batch_solver(N, 0.1, 1)
        s.set_goals(_as_f64(self.cost.xg).reshape(1, -1))
        s.set_trajectory(_as_f64(x)[None], _as_f64(u)[None])
        # alpha = 0 with dz = 0: the trial point equals (x, u)
        s.stage_dynamics()
        J, c, D = s.stage_merit(0.0)
        return float(J[0])


# --------------------------------------------------------------------------------------------------- standalone PCG
_PCG_ROBOT_BY_BLOCK = {2: "pend", 4: "arm2", 6: "arm3", 8: "arm4", 12: "arm6"}


class PCG:
    """GBD-PCG-Python's `PCG(A, b, block_size, Nblocks, guess=None, options={}).solve()` (PCG.py:5-214) on the GPU kernel.

    A must be block-tridiagonal with `Nblocks` blocks of `block_size` (the structure of the Schur complement the solver builds);
    the kernels are compiled per state dimension, so block_size must be one of 2, 4, 6, 8, 12 (the built-in robots).
    Preconditioners 'J', 'BJ', 'SS' (PCG.py:113-212); `solve()` returns `(x, (trace, trace2))` like the reference, where trace holds
    |r^T Pinv r| per iteration and trace2 (the reference's extra true-residual matvec per iteration, PCG.py:83,96) is left empty."""

    def __init__(self, A, b, block_size, Nblocks, guess=None, options=None, overloading=False):
        options = {} if options is None else options
        if overloading:
            raise ValueError("the operator-overloading tracer is not supported")
        self.A = np.asarray(A, dtype=np.float64)
        self.b = np.asarray(b, dtype=np.float64).reshape(-1)
        self.block_size, self.Nblocks = int(block_size), int(Nblocks)
        self.guess = None if guess is None else np.asarray(guess, dtype=np.float64).reshape(-1)
        self.options = options
        self.set_default_options(options)
        if self.block_size not in _PCG_ROBOT_BY_BLOCK:
            raise ValueError("block_size must be one of %r" % sorted(_PCG_ROBOT_BY_BLOCK))
        n = self.block_size * self.Nblocks
        if self.A.shape != (n, n) or self.b.shape != (n,):
            raise ValueError("A must be (block_size*Nblocks)^2 and b of matching length")
        self.Pinv = None

    def set_default_options(self, options):
        options.setdefault("exit_tolerance", 1e-6)
        options.setdefault("max_iter", 100)
        options.setdefault("DEBUG_MODE", False)
        options.setdefault("RETURN_TRACE", False)
        options.setdefault("preconditioner_type", "BJ")
        self.validate_precon_type(options["preconditioner_type"])

    def validate_precon_type(self, precon_type):
        if precon_type not in ["J", "BJ", "SS"]:
            raise ValueError("Invalid preconditioner options are [J : Jacobi, BJ: Block-Jacobi, SS: Symmetric Stair] ('0' is not offered on the GPU)")

    def update_A(self, A):
        self.A = np.asarray(A, dtype=np.float64)

    def update_b(self, b):
        self.b = np.asarray(b, dtype=np.float64).reshape(-1)

    def update_guess(self, guess):
        self.guess = np.asarray(guess, dtype=np.float64).reshape(-1)

    def update_exit_tolerance(self, tol):
        self.options["exit_tolerance"] = tol

    def update_max_iter(self, max_iter):
        self.options["max_iter"] = max_iter

    def update_preconditioner_type(self, type):
        self.validate_precon_type(type)
        self.options["preconditioner_type"] = type

    def update_DEBUG_MODE(self, mode):                  # PCG.py:45-46
        self.options["DEBUG_MODE"] = mode

    def update_RETURN_TRACE(self, mode):                # PCG.py:48-49
        self.options["RETURN_TRACE"] = mode

    def _blocks(self):
        nb, N = self.block_size, self.Nblocks
        A = self.A
        Sd = np.zeros((N, nb, nb)); So = np.zeros((N, nb, nb))
        mask = np.zeros_like(A, dtype=bool)
        for k in range(N):
            sl = slice(k * nb, (k + 1) * nb)
            Sd[k] = A[sl, sl]; mask[sl, sl] = True
            if k > 0:
                pl = slice((k - 1) * nb, k * nb)
                So[k] = A[sl, pl]; mask[sl, pl] = True; mask[pl, sl] = True
                if not np.allclose(A[pl, sl], A[sl, pl].T, rtol=1e-12, atol=1e-300):
                    raise ValueError("A must be symmetric")
        if np.any(A[~mask] != 0.0):
            raise ValueError("A must be block-tridiagonal (GBD-PCG's Schur complement structure)")
        return Sd, So

    def solve(self):
        nb, N = self.block_size, self.Nblocks
        Sd, So = self._blocks()
        b = self.b.copy()
        if self.guess is not None and np.any(self.guess != 0):
            b = b - self.A @ self.guess            # PCG from x0 == PCG from zero on the shifted system
        plant = URDFPlant(options={"path_to_urdf": _PCG_ROBOT_BY_BLOCK[nb]})
        n = plant.get_num_pos()
        cost = QuadraticCost(np.eye(2 * n), np.eye(2 * n), np.eye(n), np.zeros(2 * n))
        s = BatchSolver(plant, cost, None, N=max(N, 2), dt=0.1, batch=1)
        if N < 2:
            raise ValueError("Nblocks >= 2 required")
        method = {"J": SQPSolverMethods.PCG_J, "BJ": SQPSolverMethods.PCG_BJ, "SS": SQPSolverMethods.PCG_SS}[self.options["preconditioner_type"]]
        # The kernels factor the diagonal blocks of -S by Cholesky (the solver's Schur complement is NEGATIVE definite), while the
        # reference inverts blocks with np.linalg.inv and works for either sign (its own test.py passes a positive-definite A).
        # PCG on (-A, -b) produces exactly the iterates of PCG on (A, b) (negation is exact in floating point: r, Pinv and nu change
        # sign, alpha, beta, p, x and |nu| do not), so a positive-definite system is handed to the kernel negated.
        sign = -1.0 if float(np.einsum("kii->", Sd)) > 0.0 else 1.0
        s.set_block_system((sign * Sd).reshape(1, N, nb * nb), (sign * So).reshape(1, N, nb * nb), (sign * b).reshape(1, N, nb))
        s.stage_precond(method)
        it = s.stage_pcg(method, self.options["exit_tolerance"], self.options["max_iter"])
        x = s.fetch("l")[0].reshape(-1)
        if not np.all(np.isfinite(x)):
            raise ValueError("PCG: the diagonal blocks of A must be (positive or negative) definite")
        if self.guess is not None:
            x = x + self.guess
        trace = s.fetch("nu_trace")[0, :int(it[0]) + 1].tolist()
        self.Pinv = sign * s.fetch("Pd")[0].reshape(N, nb, nb)
        self.iterations = int(it[0])
        return x.reshape(-1, 1), (trace, [])
