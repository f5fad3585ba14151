"""Per-knot device math (csrc/b2t_core.cuh + generated model header) compiled for the host and checked against the
oracle / the reference golden vectors.  Catches arithmetic bugs without a GPU; the kernels proper are tested with -m gpu."""
import ctypes

import numpy as np
import pytest

from conftest import load_npz, relerr
from oracle import rbd, plant, cost as ocost, constraint as ocons

hostemu = pytest.importorskip("hostemu")
dp = ctypes.POINTER(ctypes.c_double)
ip = ctypes.POINTER(ctypes.c_int)


def P(a):
    return a.ctypes.data_as(dp)


def PI(a):
    return a.ctypes.data_as(ip)


@pytest.mark.parametrize("name", ["pend", "arm2", "arm3", "arm6"])
@pytest.mark.parametrize("integ", [0, 1])
def test_dynamics_math(name, integ, oracle_models):
    lib = hostemu.load(name)
    D = load_npz("dynamics.npz")
    m = oracle_models[name]
    n = m.n
    q, qd, u = D[name + "/q"], D[name + "/qd"], D[name + "/u"]
    x = np.ascontiguousarray(np.concatenate([q, qd], -1))
    u = np.ascontiguousarray(u)
    cnt = x.shape[0]
    qdd = np.zeros((cnt, n)); Minv = np.zeros((cnt, n * n)); dqdd = np.zeros((cnt, n * 3 * n)); xn = np.zeros((cnt, 2 * n))
    AB = np.zeros((cnt, 2 * n * 3 * n))
    lib.he_dynamics(cnt, P(x), P(u), ctypes.c_double(0.1), ctypes.c_double(-9.81), integ, P(qdd), P(Minv), P(dqdd), P(xn), P(AB))
    assert relerr(qdd, D[name + "/qdd"]) < 1e-13
    assert relerr(Minv.reshape(cnt, n, n), D[name + "/Minv"]) < 1e-13
    assert relerr(dqdd.reshape(cnt, n, 3 * n), D[name + "/dqdd"]) < 1e-12
    assert relerr(xn, D[name + "/xn%d" % integ]) < 1e-13
    ABr = AB.reshape(cnt, 2 * n, 3 * n)
    assert relerr(ABr[:, :, :2 * n], D[name + "/A%d" % integ]) < 1e-12
    assert relerr(ABr[:, :, 2 * n:], D[name + "/B%d" % integ]) < 1e-13


@pytest.mark.parametrize("name", ["pend", "arm2", "arm3", "arm6"])
@pytest.mark.parametrize("integ", [2, 3])
def test_multi_stage_integrator_math(name, integ, oracle_models):
    """integrator_multi_value / integrator_multi_AB (midpoint, rk3 exactly as TrajoptPlant.py:140-205 computes them) against the
    outputs of the unmodified reference (tests/golden/integrators.npz)."""
    lib = hostemu.load(name)
    D = load_npz("integrators.npz")
    n = oracle_models[name].n
    x = np.ascontiguousarray(np.concatenate([D[name + "/q"], D[name + "/qd"]], -1))
    u = np.ascontiguousarray(D[name + "/u"])
    cnt = x.shape[0]
    xn = np.zeros((cnt, 2 * n)); AB = np.zeros((cnt, 2 * n * 3 * n))
    lib.he_integrator_multi(cnt, integ, P(x), P(u), ctypes.c_double(0.1), ctypes.c_double(-9.81), P(xn), P(AB))
    assert relerr(xn, D[name + "/xn%d" % integ]) < 1e-13
    ABr = AB.reshape(cnt, 2 * n, 3 * n)
    assert relerr(ABr[:, :, :2 * n], D[name + "/A%d" % integ]) < 1e-12
    assert relerr(ABr[:, :, 2 * n:], D[name + "/B%d" % integ]) < 1e-12


@pytest.mark.parametrize("name,kind", [("arm2", 1), ("arm2", 0), ("arm6", 0), ("arm3", 1), ("arm6", 1), ("cartpole", 1)])
def test_cost_math(name, kind, oracle_models):
    """kind 1 = end-effector cost: 4 x 4 weights on (x, y, vx, vy); n = 2 is the reference's literal arithmetic, n > 2 the exact
    n-joint generalisation."""
    lib = hostemu.load(name)
    if name in oracle_models:
        m = oracle_models[name]
    else:       # not a reference model: extracted by this repository's own URDF reader
        from oracle import rbd
        from trajoptmpcreference_b200 import model as pmodel
        m = rbd.Model(pmodel.extract_model(pmodel.builtin_urdf(name)))
    n = m.n; nx = 2 * n; nm = 3 * n
    ne = 4 if kind == 1 else nx
    rng = np.random.default_rng(5)
    A = rng.uniform(-1, 1, (ne, ne)); Q = A @ A.T + np.eye(ne)
    A = rng.uniform(-1, 1, (ne, ne)); QF = 10 * (A @ A.T) + np.eye(ne)
    A = rng.uniform(-1, 1, (n, n)); R = A @ A.T + 0.1 * np.eye(n)
    xg = rng.uniform(-1, 1, ne)
    N = 7
    X = rng.uniform(-1.5, 1.5, (N, nx)); U = rng.uniform(-1, 1, (N - 1, n))
    c = ocost.UrdfCost(m, Q, QF, R, xg, QF_start=4) if kind == 1 else ocost.QuadraticCost(Q, QF, R, xg, QF_start=4)
    Upad = np.zeros((N, n)); Upad[:N - 1] = U
    val = np.zeros(N); grad = np.zeros((N, nm)); hess = np.zeros((N, nm * nm))
    kidx = np.arange(N, dtype=np.int32); term = np.zeros(N, dtype=np.int32); term[-1] = 1

    def pad(M, size):       # C-ABI layout: the packed ne x ne weights first, zeros after
        out = np.zeros(size); out[:M.size] = M.reshape(-1); return out
    lib.he_cost(N, kind, 4, P(pad(Q, nx * nx)), P(pad(QF, nx * nx)), P(np.ascontiguousarray(R)), P(pad(xg, nx)), P(X), P(Upad),
                PI(kidx), PI(term), P(val), P(grad), P(hess))
    assert relerr(val, c.values(X, U)) < 1e-13
    assert relerr(grad, c.gradients(X, U)) < 1e-13
    assert relerr(hess.reshape(N, nm, nm), c.hessians(X, U)) < 1e-13
    if kind == 1 and n > 2:     # unpinned: check the analytic gradient against central differences of the value
        eps = 1e-6
        for k in (0, N - 1):
            for i in range(nx):
                Xp = X.copy(); Xm = X.copy(); Xp[k, i] += eps; Xm[k, i] -= eps
                fd = (c.values(Xp, U)[k] - c.values(Xm, U)[k]) / (2 * eps)
                assert abs(fd - grad[k, i]) < 1e-6 * max(1.0, abs(fd))


def test_soft_math():
    lib = hostemu.load("arm3")
    n = 3; nx = 6; nm = 9; N = 6
    rng = np.random.default_rng(3)
    cons = ocons.SoftConstraints(n, n, n, N)
    cons.set_joint_limits([0.4, 0.5, 0.6], [-0.4, -0.3, -0.2], "AUGMENTED_LAGRANGIAN")
    cons.set_torque_limits([0.5], [-0.5], "QUADRATIC_PENALTY")
    for lim in cons.limits.values():
        lim.mu[:] = rng.uniform(0.5, 2, lim.mu.shape); lim.lam[:] = rng.uniform(-0.1, 0.1, lim.lam.shape)
    X = rng.uniform(-1, 1, (N, nx)); U = rng.uniform(-1, 1, (N - 1, n))
    Z = np.zeros((N, nm)); Z[:, :nx] = X; Z[:N - 1, nx:] = U
    mode = np.array([2, 2, 2, 0, 0, 0, 1, 1, 1], dtype=np.int32)
    lb = np.array([-0.4, -0.3, -0.2, 0, 0, 0, -0.5, -0.5, -0.5]); ub = np.array([0.4, 0.5, 0.6, 0, 0, 0, 0.5, 0.5, 0.5])
    mu = np.ones((N, 2 * nm)); lam = np.zeros((N, 2 * nm))
    j, tq = cons.limits["joint"], cons.limits["torque"]
    mu[:, 0:3] = j.mu[:3].T; mu[:, nm:nm + 3] = j.mu[3:].T; lam[:, 0:3] = j.lam[:3].T; lam[:, nm:nm + 3] = j.lam[3:].T
    mu[:N - 1, 6:9] = tq.mu[:3].T; mu[:N - 1, nm + 6:nm + 9] = tq.mu[3:].T
    lam[:N - 1, 6:9] = tq.lam[:3].T; lam[:N - 1, nm + 6:nm + 9] = tq.lam[3:].T
    term = np.zeros(N, dtype=np.int32); term[-1] = 1
    val = np.zeros(N); gck = np.zeros((N, nm))
    lib.he_soft(N, PI(mode), P(lb), P(ub), P(Z), P(mu), P(lam), PI(term), P(val), P(gck))
    assert relerr(val, cons.values(X, U)) < 1e-13
    assert relerr(gck, cons.gradients(X, U)) < 1e-13


def test_spd_inverse():
    lib = hostemu.load("arm2")
    rng = np.random.default_rng(1)
    A = rng.uniform(-1, 1, (12, 12)); A = A @ A.T + 0.5 * np.eye(12)
    B = np.ascontiguousarray(A.copy())
    lib.he_spd_inverse(12, P(B))
    assert relerr(B, np.linalg.inv(A)) < 1e-11


@pytest.mark.parametrize("n", [4, 12])
def test_spd_inverse_packed(n):
    lib = hostemu.load("arm2")
    rng = np.random.default_rng(2)
    A = rng.uniform(-1, 1, (n, n)); A = A @ A.T + 0.5 * np.eye(n)
    packed = np.array([A[i, j] for i in range(n) for j in range(i + 1)])
    getattr(lib, "he_spd_inverse_packed%d" % n)(P(packed))
    inv = np.linalg.inv(A)
    got = np.zeros((n, n))
    k = 0
    for i in range(n):
        for j in range(i + 1):
            got[i, j] = got[j, i] = packed[k]; k += 1
    assert relerr(got, inv) < 1e-11


@pytest.mark.parametrize("name", ["arm2", "arm3", "arm6", "cartpole"])
def test_exact_hessian_mode(name, oracle_models):
    """UrdfCost.hess_mode 1 (TrajoptCost.py:494-499; crashes in the reference, so the exact Hessian of 0.5 e^T Q e is the spec):
    the device code against the oracle's second_order_term, and -- for n > 2, where the gradient is exact -- against central
    differences of the analytic gradient (the Gauss-Newton Hessian of mode 0 does not pass that check)."""
    lib = hostemu.load(name)
    if name in oracle_models:
        m = oracle_models[name]
    else:
        from oracle import rbd
        from trajoptmpcreference_b200 import model as pmodel
        m = rbd.Model(pmodel.extract_model(pmodel.builtin_urdf(name)))
    n = m.n; nx = 2 * n; nm = 3 * n
    rng = np.random.default_rng(8)
    A = rng.uniform(-1, 1, (4, 4)); Q = A @ A.T + np.eye(4)
    QF = 10 * Q
    R = 0.1 * np.eye(n)
    xg = rng.uniform(-1, 1, 4)
    N = 5
    X = rng.uniform(-1.5, 1.5, (N, nx)); U = rng.uniform(-1, 1, (N - 1, n))
    c = ocost.UrdfCost(m, Q, QF, R, xg)
    c.hess_mode = 1
    Upad = np.zeros((N, n)); Upad[:N - 1] = U
    val = np.zeros(N); grad = np.zeros((N, nm)); hess = np.zeros((N, nm * nm))
    kidx = np.arange(N, dtype=np.int32); term = np.zeros(N, dtype=np.int32); term[-1] = 1

    def pad(M, size):
        out = np.zeros(size); out[:M.size] = M.reshape(-1); return out
    lib.he_cost_mode(N, 1, -1, 1, P(pad(Q, nx * nx)), P(pad(QF, nx * nx)), P(np.ascontiguousarray(R)), P(pad(xg, nx)), P(X), P(Upad),
                     PI(kidx), PI(term), P(val), P(grad), P(hess))
    H = hess.reshape(N, nm, nm)
    Ho = c.hessians(X, U)
    assert relerr(H, Ho) < 1e-12
    c0 = ocost.UrdfCost(m, Q, QF, R, xg)
    if name != "cartpole":      # (the cart-pole's offset point (0,1,0) lies on the pole's y axis: its planar position does not depend on the angle)
        assert relerr(Ho, c0.hessians(X, U)) > 1e-3                  # the second-order term is not negligible at a random point
    assert np.max(np.abs(H - np.swapaxes(H, 1, 2))) < 1e-12 * np.max(np.abs(H))
    if n > 2:
        eps = 1e-6
        for k in (0, N - 1):
            for i in range(nx):
                Xp = X.copy(); Xm = X.copy(); Xp[k, i] += eps; Xm[k, i] -= eps
                fd = (c.gradients(Xp, U)[k][:nx] - c.gradients(Xm, U)[k][:nx]) / (2 * eps)
                assert np.max(np.abs(fd - H[k, :nx, i])) < 1e-6 * max(1.0, np.max(np.abs(H[k])))


@pytest.mark.parametrize("name", ["pend", "arm2", "arm3", "arm6", "cartpole"])
def test_articulated_body_solve_equals_minv_times_rhs(name, oracle_models):
    """forward_dynamics_qdd (qdd = M^-1 (u - c) by the single-right-hand-side articulated-body recursion, used at the line-search trial
    points) against the oracle's Minv (u - c) (URDFPlant.forward_dynamics, TrajoptPlant.py:283-299)."""
    from oracle import rbd, plant as oplant
    lib = hostemu.load(name)
    if name in oracle_models:
        m = oracle_models[name]
    else:
        from trajoptmpcreference_b200 import model as pmodel
        m = rbd.Model(pmodel.extract_model(pmodel.builtin_urdf(name)))
    n = m.n
    rng = np.random.default_rng(12)
    cnt = 9
    X = rng.uniform(-2, 2, (cnt, 2 * n)); U = rng.uniform(-3, 3, (cnt, n))
    qdd = np.zeros((cnt, n))
    lib.he_qdd_solve(cnt, P(X), P(U), ctypes.c_double(-9.81), P(qdd))
    ref = oplant.forward_dynamics(m, X, U, -9.81)
    assert relerr(qdd, ref) < 1e-12
