"""-m gpu: every kernel stage through the C ABI against the oracle / reference golden vectors (fp64, tight tolerances)."""
import numpy as np
import pytest

import trajoptmpcreference_b200 as t
from conftest import load_npz, relerr
from gpu_common import make_pair
from oracle import kkt, sqp, dense

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["pend", "arm2", "arm3", "arm6"])
@pytest.mark.parametrize("integ", [0, 1])
def test_dynamics_kernels_vs_reference(name, integ, oracle_models):
    """k_fd + k_fd_grad against values recorded from the unmodified reference (tests/golden/dynamics.npz)."""
    D = load_npz("dynamics.npz")
    n = oracle_models[name].n
    q, qd, u = D[name + "/q"], D[name + "/qd"], D[name + "/u"]
    P = q.shape[0]
    (plant, pc, _), _ = make_pair(name, 2, oracle_models, integrator=integ, cost_kind="quadratic")
    s = t.BatchSolver(plant, pc, None, N=2, dt=0.1, batch=P)
    x = np.zeros((P, 2 * n, 2)); uu = np.zeros((P, n, 1))
    x[:, :n, 0] = q; x[:, n:, 0] = qd; uu[:, :, 0] = u
    s.set_trajectory(x, uu)
    s.stage_dynamics()
    dq = s.fetch("dqdd")[:, 0].reshape(P, n, 3 * n)
    assert relerr(dq, D[name + "/dqdd"]) < 1e-12
    assert relerr(s.fetch("xkp1")[:, 0], D[name + "/xn%d" % integ]) < 1e-13
    # reference-API callbacks of the plant object
    A, B = plant.integrator(x[0, :, 0], uu[0, :, 0], 0.1, return_gradient=True)
    assert relerr(A, D[name + "/A%d" % integ][0]) < 1e-12 and relerr(B, D[name + "/B%d" % integ][0]) < 1e-13
    assert relerr(plant.forward_dynamics(x[0, :, 0], uu[0, :, 0]), D[name + "/qdd"][0]) < 1e-12


@pytest.mark.parametrize("name", ["pend", "arm2", "arm3", "arm6"])
@pytest.mark.parametrize("integ", [2, 3])
def test_multi_stage_integrator_kernels_vs_reference(name, integ, oracle_models):
    """Integrator types 2 (midpoint) and 3 (rk3) -- k_fd's multi-stage branch and k_ab_multi -- against the step and (A, B) the
    unmodified reference returns (tests/golden/integrators.npz; TrajoptPlant.py:140-205 as written), through the batch
    workspace and through the plant object's reference-API callbacks."""
    D = load_npz("integrators.npz")
    n = oracle_models[name].n
    q, qd, u = D[name + "/q"], D[name + "/qd"], D[name + "/u"]
    P = q.shape[0]
    (plant, pc, _), _ = make_pair(name, 2, oracle_models, integrator=integ, cost_kind="quadratic")
    s = t.BatchSolver(plant, pc, None, N=2, dt=0.1, batch=P)
    x = np.zeros((P, 2 * n, 2)); uu = np.zeros((P, n, 1))
    x[:, :n, 0] = q; x[:, n:, 0] = qd; uu[:, :, 0] = u
    s.set_trajectory(x, uu)
    s.stage_dynamics()
    AB = s.fetch("AB")[:, 0].reshape(P, 2 * n, 3 * n)
    assert relerr(AB[:, :, :2 * n], D[name + "/A%d" % integ]) < 1e-12
    assert relerr(AB[:, :, 2 * n:], D[name + "/B%d" % integ]) < 1e-12
    assert relerr(s.fetch("xkp1")[:, 0], D[name + "/xn%d" % integ]) < 1e-13
    with pytest.raises(t.B2TError):
        s.fetch("dqdd")                     # stage-wise for these integrators: the workspace holds [A B] instead
    for i in (0, P - 1):
        A, B = plant.integrator(x[i, :, 0], uu[i, :, 0], 0.1, return_gradient=True)
        assert relerr(A, D[name + "/A%d" % integ][i]) < 1e-12 and relerr(B, D[name + "/B%d" % integ][i]) < 1e-12
        assert relerr(plant.integrator(x[i, :, 0], uu[i, :, 0], 0.1), D[name + "/xn%d" % integ][i]) < 1e-13
    # the per-point callbacks of such a plant still answer (Euler probe): dqdd does not depend on the integrator
    Dd = load_npz("dynamics.npz")
    assert relerr(plant.forward_dynamics_gradient(x[0, :, 0], uu[0, :, 0]), Dd[name + "/dqdd"][0]) < 1e-12


@pytest.mark.parametrize("name", ["pend", "arm3"])
def test_rk4_steps_but_its_gradient_raises_like_the_reference(name, oracle_models):
    """Integrator type 4: the reference's step works, its gradient branch raises TypeError (TrajoptPlant.py:259), so every solve does."""
    D = load_npz("integrators.npz")
    n = oracle_models[name].n
    plant = t.URDFPlant(integrator_type=4, options={"path_to_urdf": name})
    x = np.concatenate([D[name + "/q"][1], D[name + "/qd"][1]]); u = D[name + "/u"][1]
    assert relerr(plant.integrator(x, u, 0.1), D[name + "/xn4"][1]) < 1e-13
    assert str(D[name + "/rk4_gradient_raises"]) == "TypeError"
    with pytest.raises(TypeError):
        plant.integrator(x, u, 0.1, return_gradient=True)
    cost = t.QuadraticCost(np.eye(2 * n), np.eye(2 * n), np.eye(n), np.zeros(2 * n))
    with pytest.raises(TypeError):
        t.TrajoptMPCReference(plant, cost).SQP(np.zeros((2 * n, 5)), np.zeros((n, 4)), 5, 0.1, t.SQPSolverMethods.PCG_SS)


def _kkt_case(tag, oracle_models):
    K = load_npz("kkt.npz")
    robot = {"arm2_urdf": "arm2", "arm3_qc": "arm3", "arm6_qc": "arm6", "pend_al": "pend"}[tag]
    x, u, xs, xg = K[tag + "/x"], K[tag + "/u"], K[tag + "/xs"], K[tag + "/xg"]
    N = x.shape[1]
    limits = {"torque": ([0.3], [-0.3], "AUGMENTED_LAGRANGIAN")} if tag == "pend_al" else None
    (plant, pc, pcons), (m, oc, ocn) = make_pair(robot, N, oracle_models, xg=xg, limits=limits)
    if tag == "pend_al":
        ocn.limits["torque"].lam[:] = K[tag + "/lam"]; ocn.limits["torque"].mu[:] = K[tag + "/mu"]
        pcons.torque_limits.augmented_lagrangian_lambda[:] = K[tag + "/lam"]; pcons.torque_limits.quadratic_penalty_mu[:] = K[tag + "/mu"]
    return K, plant, pc, pcons, m, oc, ocn, x, u, xs, N


@pytest.mark.parametrize("tag,dense_kkt", [("arm2_urdf", False), ("arm3_qc", False), ("arm3_qc", True), ("arm6_qc", False), ("arm6_qc", True),
                                       ("pend_al", False), ("pend_al", True)])
@pytest.mark.parametrize("batch", [1, 3])
def test_kkt_schur_pcg_recover_merit(tag, dense_kkt, batch, oracle_models):
    """dense_kkt=False: structured kernels when the cost is diagonal (k_kkt_diag / k_schur_rows / k_recover_diag);
    dense_kkt=True: general dense-G_k kernels (always used for UrdfCost)."""
    K, plant, pc, pcons, m, oc, ocn, x, u, xs, N = _kkt_case(tag, oracle_models)
    n = m.n; nx = 2 * n; mm = 3 * n
    s = t.BatchSolver(plant, pc, pcons, N=N, dt=0.1, batch=batch, dense_kkt=dense_kkt)
    s.set_trajectory(np.broadcast_to(x[None], (batch,) + x.shape), np.broadcast_to(u[None], (batch,) + u.shape))
    s.set_initial_state(np.broadcast_to(xs[None], (batch, nx)))
    if pcons is not None:
        mu, lam, phi = pcons.pack(N)
        s.set_multipliers(*[np.broadcast_to(a[None], (batch,) + a.shape) for a in (mu, lam, phi)])
    rho = 1e-3
    X, U = x.T.copy(), u.T.copy()
    blocks = kkt.form_blocks(m, oc, ocn, X, U, xs, 0.1)
    sch = kkt.schur(blocks, rho, nx)
    for method, kind in ((t.SQPSolverMethods.PCG_SS, "SS"), (t.SQPSolverMethods.PCG_BJ, "BJ"), (t.SQPSolverMethods.PCG_J, "J")):
        s.stage_kkt(rho, method)
        for b in range(batch):
            assert relerr(s.fetch("g")[b], blocks["g"]) < 1e-12
            Gh = s.fetch("Ghat")[b].reshape(N, mm, mm)
            assert relerr(Gh, sch["Ghat"]) < 1e-11
            assert relerr(s.fetch("Sd")[b].reshape(N, nx, nx), sch["Sd"]) < 1e-11
            assert relerr(s.fetch("So")[b].reshape(N, nx, nx)[1:], sch["So"]) < 1e-11
            assert relerr(s.fetch("gamma")[b], sch["gamma"]) < 1e-10
        Pd, Po = kkt.preconditioner(sch["Sd"], sch["So"], kind)
        assert relerr(s.fetch("Pd")[0].reshape(N, nx, nx), Pd) < 1e-9
        # reference dense matrices as recorded from the unmodified reference
        assert relerr(dense.assemble_bt(s.fetch("Sd")[0].reshape(N, nx, nx), s.fetch("So")[0].reshape(N, nx, nx)[1:]), K[tag + "/S"]) < 1e-11
        it = s.stage_pcg(method)
        l_or, trace = kkt.pcg(sch["Sd"], sch["So"], sch["gamma"], Pd, Po)
        if kind != "J":
            # the exit test |nu| < 1e-6 is a threshold on a rounding-sensitive quantity: on this random (non-converged)
            # trajectory the count may differ by one, and then only if the oracle's |nu| sits at the threshold there
            assert len(trace) == len(K["%s/trace_%s" % (tag, kind)])
            assert len(set(it.tolist())) == 1
            dif = int(it[0]) - (len(trace) - 1)
            assert abs(dif) <= 1, (kind, it, len(trace) - 1)
            if dif != 0:
                assert trace[min(int(it[0]), len(trace) - 1)] < 1e-4, (kind, it, trace[-3:])     # |nu| drops ~50x per iteration here
            s.stage_recover()
            dz_or = kkt.recover(blocks, sch, l_or, nx)
            for b in range(batch):      # PCG stops at |r^T Pinv r| < 1e-6 (absolute): the iterate is only that accurate
                assert np.max(np.abs(s.fetch("l")[b] - l_or)) < 2e-3 * max(1.0, np.max(np.abs(l_or)))
                assert np.max(np.abs(s.fetch("dz")[b] - dz_or)) < 2e-3 * max(1.0, np.max(np.abs(dz_or)))
            # converged PCG (tolerance far below rounding): the kernel must reach the exact solution of S l = gamma
            s.stage_pcg(method, tol=1e-26, max_iter=100)
            s.stage_recover()
            l_ex = kkt.bt_solve_dense(sch["Sd"], sch["So"], sch["gamma"])
            dz_ex = kkt.recover(blocks, sch, l_ex, nx)
            assert np.max(np.abs(s.fetch("l")[0] - l_ex)) < 1e-7 * max(1.0, np.max(np.abs(l_ex)))
            assert np.max(np.abs(s.fetch("dz")[batch - 1] - dz_ex)) < 1e-7 * max(1.0, np.max(np.abs(dz_ex)))
    # exact methods S / N (block-tridiagonal factorisation) against the reference's recorded np.linalg.solve results
    for method, key in ((t.SQPSolverMethods.S, "dxul_S"), (t.SQPSolverMethods.N, "dxul_N")):
        s.stage_kkt(rho, method); it = s.stage_pcg(method); s.stage_recover()
        assert it.tolist() == [0] * batch
        dzg, lg = s.fetch("dz")[batch - 1], s.fetch("l")[batch - 1]
        packed = np.concatenate([dzg[:N - 1].reshape(-1), dzg[N - 1, :nx], lg.reshape(-1)])[:, None]
        assert relerr(packed, K[tag + "/" + key]) < 1e-7
    # merit terms of the trial point x - alpha dz (dz from the last recover = BJ): compare with the oracle on the SAME dz
    s.stage_kkt(rho, t.SQPSolverMethods.PCG_SS); s.stage_pcg(t.SQPSolverMethods.PCG_SS); s.stage_recover()
    dz = s.fetch("dz")[0]
    for alpha in (1.0, 0.25):
        J, c, D = s.stage_merit(alpha)
        Xn = X - alpha * dz[:, :nx]; Un = U - alpha * dz[:N - 1, nx:]
        J_or = sqp.total_cost(oc, ocn, Xn, Un)
        c_or = sqp.total_violation(m, Xn, Un, xs, 0.1)
        g_or = oc.gradients(Xn, Un)
        D_or = float(np.sum(g_or[:N - 1] * dz[:N - 1]) + np.sum(g_or[N - 1, :nx] * dz[N - 1, :nx]))
        if ocn is not None:
            sg = ocn.gradients(Xn, Un)
            D_or += float(np.sum(sg[:N - 1] * dz[:N - 1]) + np.sum(sg[N - 1, :nx] * dz[N - 1, :nx]))
        assert np.allclose(J, J_or, rtol=1e-12) and np.allclose(c, c_or, rtol=1e-11) and np.allclose(D, D_or, rtol=1e-10, atol=1e-12)
        assert relerr(s.fetch("xn")[0], Xn) < 1e-14


def test_urdf_cost_callbacks(oracle_models):
    """UrdfCost.value / gradient / hessian (TrajoptCost.py:402-519) evaluated by the cost kernels vs the oracle."""
    from oracle import cost as ocost
    m = oracle_models["arm2"]
    rng = np.random.default_rng(9)
    A = rng.uniform(-1, 1, (4, 4)); Q = A @ A.T + np.eye(4); QF = 50 * Q; R = np.diag([0.1, 0.3]); xg = np.array([-1.0, 1.5, 0.1, -0.2])
    plant = t.URDFPlant(options={"path_to_urdf": "arm2"})
    pc = t.UrdfCost(plant, Q, QF, R, xg, QF_start=3)
    oc = ocost.UrdfCost(m, Q, QF, R, xg, QF_start=3)
    for k in (1, 4):
        x = rng.uniform(-1, 1, 4); u = rng.uniform(-1, 1, 2)
        X = np.stack([x, x]); U = u[None]
        vals = oc.values(X, U); grads = oc.gradients(X, U); hess = oc.hessians(X, U)
        # oracle arrays are for knots (0 running, 1 terminal); evaluate the running knot at timestep k explicitly
        oc_k = ocost.UrdfCost(m, Q if k < 3 else QF, QF, R, xg)
        v_run = oc_k.values(X, U)[0]; g_run = oc_k.gradients(X, U)[0]; h_run = oc_k.hessians(X, U)[0]
        assert abs(pc.value(x, u, k) - v_run) < 1e-12 * max(1, abs(v_run))
        assert relerr(pc.gradient(x, u, k), g_run) < 1e-12
        assert relerr(pc.hessian(x, u, k), h_run) < 1e-12
        assert abs(pc.value(x, None, 9) - vals[1]) < 1e-12 * max(1, abs(vals[1]))
        assert relerr(pc.gradient(x, None, 9), grads[1][:4]) < 1e-12
        assert relerr(pc.hessian(x, None, 9), hess[1][:4, :4]) < 1e-12


@pytest.mark.parametrize("nb,N,kind", [(4, 10, "SS"), (12, 8, "BJ"), (6, 5, "J"), (12, 64, "SS")])
def test_standalone_pcg_class(nb, N, kind):
    """The reference's own PCG test idea (GBD-PCG-Python/test.py: PSD block system vs np.linalg.solve) through the drop-in PCG class,
    plus the oracle's trace."""
    rng = np.random.default_rng(nb * 100 + N)
    # negative-definite block-tridiagonal matrix like the Schur complement: S = -(L L^T), L block lower-bidiagonal
    L = np.zeros((nb * N, nb * N))
    for k in range(N):
        L[k * nb:(k + 1) * nb, k * nb:(k + 1) * nb] = rng.uniform(-1, 1, (nb, nb)) + 3 * np.eye(nb)
        if k > 0:
            L[k * nb:(k + 1) * nb, (k - 1) * nb:k * nb] = 0.5 * rng.uniform(-1, 1, (nb, nb))
    S = -(L @ L.T)
    b = rng.uniform(-1, 1, nb * N)
    pcg = t.PCG(S, b, nb, N, options={"preconditioner_type": kind, "exit_tolerance": 1e-6, "max_iter": 100})
    x, (trace, _) = pcg.solve()
    Sd = np.stack([S[k * nb:(k + 1) * nb, k * nb:(k + 1) * nb] for k in range(N)])
    So = np.stack([S[k * nb:(k + 1) * nb, (k - 1) * nb:k * nb] for k in range(1, N)])
    Pd, Po = kkt.preconditioner(Sd, So, kind)
    lo, tr_o = kkt.pcg(Sd, So, b.reshape(N, nb), Pd, Po)
    assert abs(len(trace) - len(tr_o)) <= (1 if kind == "J" else 0)
    k = min(len(trace), len(tr_o), 6)
    assert np.allclose(trace[:k], tr_o[:k], rtol=1e-6)
    exact = np.linalg.solve(S, b)
    assert np.max(np.abs(x[:, 0] - exact)) < 5e-3 * max(1.0, np.max(np.abs(exact)))
    # warm start (PCG.update_guess, PCG.py:33): starting at the exact solution needs no iteration to meet the tolerance
    pcg.update_guess(exact); pcg.update_exit_tolerance(1e-10)
    x2, (tr2, _) = pcg.solve()
    assert np.max(np.abs(x2[:, 0] - exact)) < 1e-8 and len(tr2) <= 3


@pytest.mark.parametrize("nb,N,kind", [(4, 4, "SS"), (12, 8, "BJ"), (6, 5, "SS"), (4, 6, "J")])
def test_standalone_pcg_positive_definite(nb, N, kind):
    """GBD-PCG-Python/test.py passes a POSITIVE-definite A ('PCG expects a positive semidefinite matrix': A = M M^T, M block
    tridiagonal-free block diagonal with symmetric blocks, utils.py:60-79); the reference inverts blocks with np.linalg.inv and so
    accepts either sign.  Here an SPD block-tridiagonal system: the result must equal np.linalg.solve and the trace / iterates must
    be those of the negated (negative-definite) system the solver itself produces."""
    rng = np.random.default_rng(nb * 10 + N)
    L = np.zeros((nb * N, nb * N))
    for k in range(N):
        blk = rng.uniform(0, 1, (nb, nb)); blk = 0.5 * (blk + blk.T) + 0.1 * np.eye(nb) + 2 * np.eye(nb)
        L[k * nb:(k + 1) * nb, k * nb:(k + 1) * nb] = blk
        if k > 0:
            L[k * nb:(k + 1) * nb, (k - 1) * nb:k * nb] = 0.3 * rng.uniform(-1, 1, (nb, nb))
    A = L @ L.T
    b = rng.uniform(0, 1, nb * N)
    pcg = t.PCG(A, b, nb, N, options={"preconditioner_type": kind})
    pcg.update_RETURN_TRACE(True); pcg.update_DEBUG_MODE(False)
    x, (trace, _) = pcg.solve()
    exact = np.linalg.solve(A, b)
    assert np.all(np.isfinite(x)) and np.max(np.abs(x[:, 0] - exact)) < 5e-3 * max(1.0, np.max(np.abs(exact)))
    neg = t.PCG(-A, -b, nb, N, options={"preconditioner_type": kind})
    xn, (trn, _) = neg.solve()
    assert np.array_equal(x, xn) and trace == trn                  # negation is exact: identical iterates
    assert np.allclose(pcg.Pinv, -neg.Pinv)
    if kind != "J":                                              # block inverse with the reference's sign
        assert np.allclose(pcg.Pinv[0], np.linalg.inv(A[:nb, :nb]), rtol=1e-9, atol=1e-12)
    pcg.update_exit_tolerance(1e-12); pcg.update_max_iter(100)
    x3, _ = pcg.solve()
    assert np.max(np.abs(x3[:, 0] - exact)) < 1e-6 * max(1.0, np.max(np.abs(exact)))


def test_soft_constraint_callbacks(oracle_models):
    """TrajoptConstraint.value_soft_constraints / jacobian_soft_constraints (TrajoptConstraint.py:295-340) through the constraint kernels,
    with non-trivial multipliers, against the oracle's element-wise restatement."""
    N = 6
    limits = {"torque": ([0.4], [-0.4], "AUGMENTED_LAGRANGIAN"), "joint": ([0.45], [-0.45], "QUADRATIC_PENALTY"),
              "velocity": ([0.3, 0.5, 0.7], [-0.3, -0.5, -0.7], "AUGMENTED_LAGRANGIAN")}
    (plant, pc, pcons), (m, oc, ocn) = make_pair("arm3", N, oracle_models, limits=limits)
    n = m.n
    rng = np.random.default_rng(3)
    for p_lim, o_lim in ((pcons.joint_limits, ocn.limits["joint"]), (pcons.velocity_limits, ocn.limits["velocity"]), (pcons.torque_limits, ocn.limits["torque"])):
        mu = rng.uniform(0.5, 2.0, p_lim.quadratic_penalty_mu.shape); lam = rng.uniform(-0.05, 0.05, mu.shape)
        p_lim.quadratic_penalty_mu[:] = mu; p_lim.augmented_lagrangian_lambda[:] = lam
        o_lim.mu[:] = mu; o_lim.lam[:] = lam
    solver = t.TrajoptMPCReference(plant, pc, pcons)           # binds the constraint object to the plant's library
    X = rng.uniform(-0.8, 0.8, (N, 2 * n)); U = rng.uniform(-0.8, 0.8, (N - 1, n))
    vals = ocn.values(X, U); grads = ocn.gradients(X, U)
    for k in (0, 2, N - 1):
        uk = U[k] if k < N - 1 else None
        assert abs(pcons.value_soft_constraints(X[k], uk, k) - vals[k]) < 1e-13 * max(1.0, abs(vals[k]))
        g = pcons.jacobian_soft_constraints(X[k], uk, k)
        assert g.shape == (3 * n, 1) and np.allclose(g[:, 0], grads[k], rtol=1e-13, atol=1e-14)


@pytest.mark.parametrize("name,N", [("arm6", 3), ("arm6", 7), ("arm6", 33), ("arm6", 64), ("arm6", 96), ("arm6", 130), ("arm2", 5), ("arm2", 200),
                                    ("arm3", 9), ("arm3", 150)])
def test_pcg_variants_over_horizon_lengths(name, N, oracle_models):
    """Every horizon length selects a PCG kernel (matrix-free k_pcg3 up to 64 knots when nx % 4 == 0, explicit k_pcg2 / k_pcg beyond or
    for other nx): the CONVERGED PCG solution must equal the exact block-tridiagonal solve (method S) of the same system, ragged sizes
    included; the two paths share nothing but the KKT blocks."""
    (plant, pc, _), (m, _, _) = make_pair(name, N, oracle_models, cost_kind="quadratic")
    n = m.n
    B = 3
    rng = np.random.default_rng(N)
    s = t.BatchSolver(plant, pc, None, N=N, dt=0.1, batch=B)
    s.set_goals(np.tile(np.concatenate([np.linspace(0.5, -0.5, n), np.zeros(n)]), (B, 1)))
    x = rng.uniform(-0.3, 0.3, (B, 2 * n, N)); u = rng.uniform(-0.3, 0.3, (B, n, N - 1))
    s.set_trajectory(x, u)
    s.set_initial_state(x[:, :, 0] + 0.01)
    s.stage_dynamics()
    s.stage_kkt(1e-3, t.SQPSolverMethods.PCG_SS)
    it = s.stage_pcg(t.SQPSolverMethods.PCG_SS, tol=1e-24, max_iter=4000)
    l_pcg = s.fetch("l")
    s.stage_kkt(1e-3, t.SQPSolverMethods.S)
    s.stage_pcg(t.SQPSolverMethods.S)
    l_s = s.fetch("l")
    assert np.all(it < 4000)
    assert np.max(np.abs(l_pcg - l_s)) < 1e-6 * np.max(np.abs(l_s))
    s.close()
