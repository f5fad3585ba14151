"""-m gpu: complete SQP solves through the drop-in API against the reference's recorded results and the oracle."""
import numpy as np
import pytest

import trajoptmpcreference_b200 as t
from conftest import load_npz
from gpu_common import make_pair, solve_meta
from oracle import sqp

pytestmark = pytest.mark.gpu

PCG_TAGS = list(solve_meta().keys())       # PCG-J/BJ/SS, the exact methods N, S, and the hard ACTIVE_SET cases of N / S
# pend_N20_S_as01: the reference's own result is not reproducible -- every 1-ulp perturbation of its linear solve changes its iteration
# counts (tests/golden/floor.json: counts_identical 0 of 4; 3 of 4 perturbed runs land on the method-N answer J = 61.41085)
UNSTABLE_IN_REFERENCE = {"pend_N20_S_as01"}


def test_hard_constraint_modes():
    """FULL_SET is refused (singular KKT in the reference); ACTIVE_SET runs with the exact methods and is refused with PCG (the
    reference hands PCG a Schur complement whose size no longer matches block_size * Nblocks)."""
    with pytest.raises(ValueError):
        t.TrajoptConstraint(1, 1, 1, 20).set_torque_limits([0.1], [-0.1], "FULL_SET", {})
    plant = t.URDFPlant(options={"path_to_urdf": "pend"})
    cost = t.QuadraticCost(np.eye(2), 100.0 * np.eye(2), 0.1 * np.eye(1), np.array([3.14159, 0.0]))
    cons = t.TrajoptConstraint(1, 1, 1, 20)
    cons.set_torque_limits([0.1], [-0.1], "ACTIVE_SET", {})
    solver = t.TrajoptMPCReference(plant, cost, cons)
    with pytest.raises(Exception, match="exact methods"):
        solver.SQP(np.zeros((2, 20)), np.zeros((1, 19)), 20, 0.1, t.SQPSolverMethods.PCG_SS, options={})
    # both exact methods give the KKT solution: the known answer of SURVEY.md section 2 (4 QP solves, J = 61.41085)
    for method in (t.SQPSolverMethods.N, t.SQPSolverMethods.S):
        x, u, e1, e2, outer, it = solver.SQP(np.zeros((2, 20)), np.zeros((1, 19)), 20, 0.1, method, options={"expected_reduction_min_SQP_DDP": -100})
        assert (e1, it) == (1, 3) and abs(solver.last_result.J[0] - 61.41084999081366) < 1e-9 * 61.4
        assert np.max(np.abs(u)) <= 0.1 + 1e-12


def _floor():
    """Measured parity floor of the reference itself (tests/golden/floor.json, scripts/parity_floor.py): how far the UNMODIFIED
    reference's final J, x, u move when S is perturbed by 1 ulp in every QP solve."""
    import json
    import os
    from conftest import GOLDEN
    with open(os.path.join(GOLDEN, "floor.json")) as f:
        return json.load(f)["cases"]


FLOOR_FACTOR = 10.0      # allowed multiple of the reference's own 1-ulp sensitivity
NORTH_STAR = 1e-9        # BASELINE.json north_star: <= 1e-9 relative in fp64 (never asserted tighter than this)


@pytest.mark.parametrize("tag", PCG_TAGS)
def test_sqp_vs_reference_golden(tag, oracle_models):
    """Same call as the reference's examples, against outputs of the unmodified reference (tests/golden/solve.npz): exits, every
    iteration count, the alpha / rho sequences identical; J, x, u, multipliers within max(1e-9 relative, 10 x the reference's own
    sensitivity to a 1-ulp perturbation of S) -- including the soft-limit cases with 165 / 255 QP solves over 10 outer iterations."""
    S = load_npz("solve.npz")
    mt = solve_meta()[tag]
    fl = _floor()[tag]
    N = mt["N"]
    (plant, pc, pcons), _ = make_pair(mt["robot"], N, oracle_models, xg=S[tag + "/xg"], limits=mt["limits"], integrator=mt["integrator"])
    solver = t.TrajoptMPCReference(plant, pc, pcons) if pcons is not None else t.TrajoptMPCReference(plant, pc)
    n = plant.get_num_pos()
    opts = dict(mt["options"]); opts["overloading"] = False
    x, u, e1, e2, outer, it = solver.SQP(np.zeros((2 * n, N)), np.zeros((n, N - 1)), N, 0.1, getattr(t.SQPSolverMethods, mt["method"]), options=opts)
    if tag in UNSTABLE_IN_REFERENCE:
        pytest.skip("the reference's recorded result for this case does not survive a 1-ulp perturbation of its own linear solve")
    assert [e1, e2, outer, it] == S[tag + "/exits"].tolist()
    rows = solver.trace[1:]            # the reference keeps the rows of the LAST outer iteration (:555)
    k = len(rows)
    if mt["method"].startswith("PCG"):
        assert [r["pcg_iters"] for r in rows] == S[tag + "/pcg_iters"].tolist()[-k:]
        # every QP solve of every outer iteration: totals over the whole solve
        assert int(solver.last_result.total_qp[0]) == len(S[tag + "/pcg_iters"]) and int(solver.last_result.total_pcg[0]) == int(S[tag + "/pcg_iters"].sum())
    else:
        assert all(r["pcg_iters"] == 0 for r in rows)
    assert int(solver.last_result.total_trials[0]) == int((S[tag + "/tr_ls"] + 1).sum())
    assert [r["line_search_iteration"] for r in rows] == S[tag + "/tr_ls"].tolist()[-k:]
    assert np.array_equal([r["alpha"] for r in rows], S[tag + "/tr_alpha"][-k:])
    assert np.allclose([r["rho"] for r in rows], S[tag + "/tr_rho"][-k:], rtol=1e-15)
    tolJ = max(FLOOR_FACTOR * fl["rel_J"], NORTH_STAR)
    tolx = max(FLOOR_FACTOR * fl["abs_x"], NORTH_STAR * max(1.0, float(np.max(np.abs(S[tag + "/x"])))))
    tolu = max(FLOOR_FACTOR * fl["abs_u"], NORTH_STAR * max(1.0, float(np.max(np.abs(S[tag + "/u"])))))
    eJ = abs(solver.last_result.J[0] - float(S[tag + "/J"])) / abs(float(S[tag + "/J"]))
    ex = float(np.max(np.abs(x - S[tag + "/x"]))); eu = float(np.max(np.abs(u - S[tag + "/u"])))
    print("%s: rel J %.1e (tol %.1e)  |dx| %.1e (tol %.1e)  |du| %.1e (tol %.1e)" % (tag, eJ, tolJ, ex, tolx, eu, tolu))
    assert eJ <= tolJ and ex <= tolx and eu <= tolu
    assert np.allclose([r["J"] for r in rows], S[tag + "/tr_J"][-k:], rtol=tolJ, atol=0)
    if pcons is not None and pcons.torque_limits.is_soft_constraint_mode():
        assert np.allclose(pcons.torque_limits.quadratic_penalty_mu, S[tag + "/mu"], rtol=1e-15)
        assert np.max(np.abs(pcons.torque_limits.augmented_lagrangian_lambda - S[tag + "/lam"])) <= NORTH_STAR * max(1.0, float(np.max(np.abs(S[tag + "/lam"]))))


@pytest.mark.parametrize("run", ["4", "3"])
def test_recorded_author_runs(run, oracle_models):
    """data/4 (= examples/twolinks.py) and data/3 as recorded by the reference's authors."""
    D = load_npz("ref_data.npz")
    N = 10
    (plant, pc, _), _ = make_pair("arm2", N, oracle_models, xg=D[run + "/xg"])
    solver = t.TrajoptMPCReference(plant, pc)
    x, u, e1, e2, outer, it = solver.SQP(np.zeros((4, N)), np.zeros((2, N - 1)), N, 0.1, t.SQPSolverMethods.PCG_SS,
                                         options={"expected_reduction_min_SQP_DDP": -100, "overloading": False})
    assert [e1, e2, outer, it] == D[run + "/exits"].tolist()
    assert solver.pcg_iters == D[run + "/pcg_iters"].tolist()
    rows = solver.trace[1:]
    assert [r["line_search_iteration"] for r in rows] == D[run + "/tr_ls"].tolist()
    assert np.array_equal([r["alpha"] for r in rows], D[run + "/tr_alpha"])
    # recorded on the authors' machine (another numpy / BLAS build): 10 x the measured 1-ulp floor of the same case
    fl = _floor()["arm2_N10_SS" if run == "4" else "arm2_N10_SS_xg3"]
    ex, eu = float(np.max(np.abs(x - D[run + "/final_x"]))), float(np.max(np.abs(u - D[run + "/final_u"])))
    eJ = float(np.max(np.abs(np.array([r["J"] for r in rows]) / D[run + "/tr_J"] - 1.0)))
    print("data/%s: rel J %.1e (floor %.1e)  |dx| %.1e (floor %.1e)  |du| %.1e (floor %.1e)" % (run, eJ, fl["rel_J"], ex, fl["abs_x"], eu, fl["abs_u"]))
    assert eJ <= FLOOR_FACTOR * fl["rel_J"] and ex <= FLOOR_FACTOR * fl["abs_x"] and eu <= FLOOR_FACTOR * fl["abs_u"]


def _batch_goals(n, B, seed):
    rng = np.random.default_rng(seed)
    xg = np.zeros((B, 2 * n))
    xg[:, :n] = rng.uniform(-0.5, 0.5, (B, n))
    return xg


@pytest.mark.parametrize("name,N,B,limits,integ", [
    ("arm6", 16, 24, None, 0),
    ("arm3", 12, 16, {"torque": ([0.4], [-0.4], "QUADRATIC_PENALTY"), "joint": ([0.45], [-0.45], "AUGMENTED_LAGRANGIAN")}, 0),
    ("arm6", 16, 12, {"torque": ([1.0], [-1.0], "QUADRATIC_PENALTY"), "joint": ([0.45], [-0.45], "QUADRATIC_PENALTY")}, 0),
    ("arm6", 12, 8, {"torque": ([1.0], [-1.0], "QUADRATIC_PENALTY"), "joint": ([0.45], [-0.45], "QUADRATIC_PENALTY")}, 2),      # midpoint
    ("arm3", 12, 8, {"torque": ([0.4], [-0.4], "AUGMENTED_LAGRANGIAN")}, 3),                                                    # rk3
])
def test_batch_vs_oracle(name, N, B, limits, integ, oracle_models):
    """Independent instances with different goals in one batch; each compared with the oracle run on its own
    (multi-coordinate box limits are UNPINNED in the reference: the oracle's element-wise restatement is the spec).
    integ 2 / 3: the reference's midpoint / rk3 integrators (general kernels, [A B] stored in full)."""
    (plant, pc, pcons), (m, oc, ocn) = make_pair(name, N, oracle_models, limits=limits, integrator=integ)
    n = m.n
    xg = _batch_goals(n, B, 11)
    solver = t.TrajoptMPCReference(plant, pc, pcons) if pcons is not None else t.TrajoptMPCReference(plant, pc)
    opts = {"expected_reduction_min_SQP_DDP": -100, "max_iter_softConstraints": 4}
    r = solver.solve_batch(np.zeros((B, 2 * n, N)), np.zeros((B, n, N - 1)), xg, N, 0.1, t.SQPSolverMethods.PCG_SS, dict(opts))
    import copy
    import sys
    import os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    from make_c4_fixture import _perturber          # S -> S (1 +/- 2^-52): the oracle's own sensitivity = the tolerance scale
    same_iters, worst = 0, (0.0, 0.0)
    for b in range(B):
        oc_b = copy.copy(oc); oc_b.xg = xg[b]
        ro = sqp.sqp(m, oc_b, copy.deepcopy(ocn), np.zeros((2 * n, N)), np.zeros((n, N - 1)), N, 0.1, "PCG-SS", dict(opts), integrator_type=integ)
        rp = sqp.sqp(m, oc_b, copy.deepcopy(ocn), np.zeros((2 * n, N)), np.zeros((n, N - 1)), N, 0.1, "PCG-SS", dict(opts, _perturb_S=_perturber(b)),
                     integrator_type=integ)
        stable = rp["pcg_iters"] == ro["pcg_iters"] and rp["ls_trials"] == ro["ls_trials"]
        fJ = abs(rp["J"] - ro["J"]) / max(1.0, abs(ro["J"])); fx = float(np.max(np.abs(rp["x"] - ro["x"])))
        same = (ro["exit_sqp"], ro["exit_soft"], ro["outer_iter"], ro["sqp_iter"]) == (r.exit_sqp[b], r.exit_soft[b], r.outer_iter[b], r.sqp_iter[b]) \
            and sum(ro["pcg_iters"]) == r.total_pcg[b] and sum(ro["ls_trials"]) == r.total_trials[b]
        same_iters += int(same)
        if same and stable:        # within 100 x the oracle's own 1-ulp sensitivity on this instance (one perturbation seed), never tighter than 1e-9
            eJ = abs(ro["J"] - r.J[b]) / max(1.0, abs(ro["J"])); ex = float(np.max(np.abs(ro["x"] - r.x[b])))
            worst = (max(worst[0], eJ), max(worst[1], ex))
            assert eJ <= max(100 * fJ, 1e-9), (b, eJ, fJ)
            assert ex <= max(100 * fx, 1e-9), (b, ex, fx)
    print("%s N=%d: %d of %d instances reproduce every oracle count; worst rel J %.1e, worst |dx| %.1e" % (name, N, same_iters, B, worst[0], worst[1]))
    assert same_iters >= int(0.9 * B), "only %d of %d instances reproduce the oracle's iteration counts" % (same_iters, B)


@pytest.mark.parametrize("name,N,B,method,iters", [("arm3", 12, 6, "S", 12), ("arm6", 16, 4, "S", 4), ("arm3", 12, 6, "PCG_SS", 12)])
def test_end_effector_cost_n_link(name, N, B, method, iters, oracle_models):
    """UrdfCost on an n > 2 planar chain (UNPINNED: the oracle's exact generalisation is the spec):
    per-knot value / gradient / Gauss-Newton Hessian through the cost callbacks, then complete solves against the oracle."""
    import copy
    m = oracle_models[name]
    n = m.n
    rng = np.random.default_rng(21)
    ang = rng.uniform(0.3, 2.8, B); rad = rng.uniform(0.5, 0.9 * n, B)
    xg = np.stack([rad * np.cos(ang), rad * np.sin(ang), np.zeros(B), np.zeros(B)], axis=1)
    (plant, pc, _), (_, oc, _) = make_pair(name, N, oracle_models, xg=xg[0], cost_kind="urdf")
    xk = rng.uniform(-1, 1, 2 * n); uk = rng.uniform(-1, 1, n)
    X = np.stack([xk, xk]); U = uk[None]
    assert abs(pc.value(xk, uk) - oc.values(X, U)[0]) < 1e-12 * max(1.0, abs(oc.values(X, U)[0]))
    assert np.allclose(pc.gradient(xk, uk), oc.gradients(X, U)[0], rtol=1e-12, atol=1e-12)
    assert np.allclose(pc.hessian(xk, uk), oc.hessians(X, U)[0], rtol=1e-12, atol=1e-12)
    assert np.allclose(pc.gradient(xk), oc.gradients(X, U)[1][:2 * n], rtol=1e-12, atol=1e-12)
    assert np.allclose(pc.delta_x(xk), oc.state_error(xk[None])[0], rtol=1e-12, atol=1e-13)
    solver = t.TrajoptMPCReference(plant, pc)
    opts = {"expected_reduction_min_SQP_DDP": -100, "max_iter_SQP_DDP": iters}
    r = solver.solve_batch(np.zeros((B, 2 * n, N)), np.zeros((B, n, N - 1)), xg, N, 0.1, getattr(t.SQPSolverMethods, method), dict(opts))
    same_iters = 0
    for b in range(B):
        oc_b = copy.copy(oc); oc_b.xg = xg[b]
        ro = sqp.sqp(m, oc_b, None, np.zeros((2 * n, N)), np.zeros((n, N - 1)), N, 0.1, method.replace("_", "-"), dict(opts))
        same = (ro["exit_sqp"], ro["sqp_iter"]) == (r.exit_sqp[b], r.sqp_iter[b]) and sum(ro["ls_trials"]) == r.total_trials[b]
        same_iters += int(same)
        if same:    # ill-conditioned (G_k + rho I has condition ~1e7) and non-convex: the two exact solvers differ by ~2e-9 in x after ONE
            # iteration and the difference grows ~10x per iteration (scripts/probe_ee.py), hence the iteration cap for arm6;
            # with PCG the 1e-6 exit test is what gets amplified
            assert abs(ro["J"] - r.J[b]) < (1e-6 if method == "S" else 1e-4) * max(1.0, abs(ro["J"]))
            assert np.max(np.abs(ro["x"] - r.x[b])) < (1e-4 if method == "S" else 1e-2)
    # exact linear solves reproduce the oracle's control flow; with PCG the Gauss-Newton Hessian of a 4-dimensional task-space error is
    # rank-deficient in the 2n-dimensional state, S is ill-conditioned, PCG stops at its 100-iteration cap and rounding decides the path
    need = B - 1 if method == "S" else B // 2
    assert same_iters >= need, "only %d of %d instances reproduce the oracle's iteration counts" % (same_iters, B)
    for b in range(B):      # the cost went down from the start point
        oc_b = copy.copy(oc); oc_b.xg = xg[b]
        assert r.J[b] < oc_b.values(np.zeros((N, 2 * n)), np.zeros((N - 1, n))).sum()


def test_full_size_properties(oracle_models):
    """BASELINE config shape (arm6, N=64, penalty box limits) at a reduced batch: size-independent properties."""
    N, B = 64, 96
    limits = {"torque": ([1.0], [-1.0], "QUADRATIC_PENALTY"), "joint": ([0.45], [-0.45], "QUADRATIC_PENALTY")}
    (plant, pc, pcons), _ = make_pair("arm6", N, oracle_models, limits=limits)
    n = 6
    rng = np.random.default_rng(1)
    xg = np.zeros((B, 12)); xg[:, :6] = rng.uniform(-0.5, 0.5, (B, 6))
    opts = {"expected_reduction_min_SQP_DDP": -100}
    solver = t.TrajoptMPCReference(plant, pc, pcons)
    x0 = np.zeros((B, 12, N)); u0 = np.zeros((B, 6, N - 1))
    r1 = solver.solve_batch(x0, u0, xg, N, 0.1, t.SQPSolverMethods.PCG_SS, dict(opts))
    r1 = {k: np.array(v) for k, v in r1.items()}
    # (a) run-to-run determinism: bitwise
    r2 = solver.solve_batch(x0, u0, xg, N, 0.1, t.SQPSolverMethods.PCG_SS, dict(opts))
    assert np.array_equal(r1["x"], r2.x) and np.array_equal(r1["u"], r2.u) and np.array_equal(r1["total_pcg"], r2.total_pcg)
    # (b) an instance's result does not depend on its position in the batch or on its neighbours
    perm = rng.permutation(B)
    r3 = solver.solve_batch(x0, u0, xg[perm], N, 0.1, t.SQPSolverMethods.PCG_SS, dict(opts))
    assert np.array_equal(r1["x"][perm], r3.x) and np.array_equal(r1["sqp_iter"][perm], r3.sqp_iter)
    # (c) every instance terminated with a legal exit code; the start point is kept (x_0 = xs up to the PCG tolerance)
    assert set(np.unique(r1["exit_sqp"])) <= {1, 2, 3} and set(np.unique(r1["exit_soft"])) <= {1, 2, 3}
    assert np.max(np.abs(r1["x"][:, :, 0])) < 1e-3
    # (d) reported J, c equal a fresh evaluation of the merit kernel at the returned trajectory
    s = solver.batch_solver(N, 0.1, B)
    s.stage_dynamics()
    J, c, _ = s.stage_merit(0.0)
    assert np.allclose(J, r3.J, rtol=1e-12) and np.allclose(c, r3.c, rtol=1e-9, atol=1e-12)
    # (e) the unconstrained anchor instance of BASELINE.md inside a batch: 6 SQP iterations, 563 PCG iterations, 16 trials
    solver2 = t.TrajoptMPCReference(plant, pc)
    xg2 = xg.copy(); xg2[5, :6] = np.linspace(0.5, -0.5, 6)
    r4 = solver2.solve_batch(x0, u0, xg2, N, 0.1, t.SQPSolverMethods.PCG_SS, dict(opts))
    assert (r4.sqp_iter[5], r4.total_pcg[5], r4.total_trials[5]) == (6, 563, 16)
    assert abs(r4.J[5] - 6.50929423656) < 1e-7       # parity floor: the reference's own J moves by ~1e-8 under 1-ulp perturbations


def test_fp32_mode(oracle_models):
    """dtype='f32' instantiates the whole path in single precision.  Measured (scripts/fp32_probe.py):
    on arm2 it tracks fp64 to ~1e-4; on arm6 the SQP iteration amplifies fp32 rounding (Ghat has entries ~1/rho = 1e3, the
    PCG exit test is an absolute 1e-6) and the iterates take a different path, so only sanity is asserted there."""
    N = 10
    (plant, pc, _), _ = make_pair("arm2", N, oracle_models, cost_kind="quadratic", xg=np.array([0.4, -0.3, 0.0, 0.0]))
    B = 8
    xg = _batch_goals(2, B, 5)
    solver = t.TrajoptMPCReference(plant, pc)
    opts = {"expected_reduction_min_SQP_DDP": -100}
    x0 = np.zeros((B, 4, N)); u0 = np.zeros((B, 2, N - 1))
    r64 = solver.solve_batch(x0, u0, xg, N, 0.1, t.SQPSolverMethods.PCG_SS, dict(opts), dtype="f64")
    r64 = {k: np.array(v) for k, v in r64.items()}
    r32 = solver.solve_batch(x0, u0, xg, N, 0.1, t.SQPSolverMethods.PCG_SS, dict(opts), dtype="f32")
    assert set(np.unique(r32.exit_sqp)) <= {1, 2, 3}
    assert np.all(np.abs(r32.J - r64["J"]) <= 6e-4 * np.maximum(1.0, np.abs(r64["J"])))
    assert np.max(np.abs(r32.x - r64["x"])) < 5e-3
    (plant6, pc6, _), _ = make_pair("arm6", 16, oracle_models)
    s6 = t.TrajoptMPCReference(plant6, pc6)
    r = s6.solve_batch(np.zeros((B, 12, 16)), np.zeros((B, 6, 15)), _batch_goals(6, B, 5), 16, 0.1, t.SQPSolverMethods.PCG_SS, dict(opts), dtype="f32")
    assert set(np.unique(r.exit_sqp)) <= {1, 2, 3} and np.all(np.isfinite(r.J)) and np.all(np.isfinite(r.x))


@pytest.mark.parametrize("name,N", [("arm2", 5), ("arm2", 33), ("arm2", 130), ("arm3", 47), ("arm6", 3), ("pend", 2)])
def test_ragged_horizons_vs_oracle(name, N, oracle_models):
    """Horizon lengths that are not multiples of the warp / block sizes (and the degenerate N = 2, 3)."""
    (plant, pc, _), (m, oc, _) = make_pair(name, N, oracle_models, cost_kind="quadratic",
                                           xg=np.array([0.4, -0.3, 0.0, 0.0]) if name == "arm2" else None)
    n = m.n
    B = 3
    xg = np.tile(np.asarray(oc.xg, dtype=float), (B, 1)); xg[1:, :n] *= np.array([[0.5], [-0.7]])
    solver = t.TrajoptMPCReference(plant, pc)
    opts = {"expected_reduction_min_SQP_DDP": -100, "max_iter_SQP_DDP": 30}
    for method, om in ((t.SQPSolverMethods.PCG_SS, "PCG-SS"), (t.SQPSolverMethods.S, "S")):
        r = solver.solve_batch(np.zeros((B, 2 * n, N)), np.zeros((B, n, N - 1)), xg, N, 0.1, method, dict(opts))
        same = 0
        for b in range(B):
            import copy
            ocb = copy.copy(oc); ocb.xg = xg[b]
            ro = sqp.sqp(m, ocb, None, np.zeros((2 * n, N)), np.zeros((n, N - 1)), N, 0.1, om, dict(opts))
            ok = (ro["exit_sqp"], ro["sqp_iter"], sum(ro["pcg_iters"]), sum(ro["ls_trials"])) == (r.exit_sqp[b], r.sqp_iter[b], r.total_pcg[b], r.total_trials[b])
            same += int(ok)
            # J always agrees; the iteration at which `delta_J < 1e-6` fires may differ by one on slowly (linearly) converging
            # cases such as arm6 / N=3, where consecutive J differ by ~1e-6 and the two solvers differ by ~1e-8
            assert abs(ro["J"] - r.J[b]) < 1e-6 * max(1.0, abs(ro["J"]))
            if ok:
                assert np.max(np.abs(ro["x"] - r.x[b])) < 1e-4
            else:
                assert abs(ro["sqp_iter"] - r.sqp_iter[b]) <= 1
        assert same >= B - 2, (name, N, om, same)


def test_in_place_mutators_invalidate_cached_workspace(oracle_models):
    """The reference API mutates costs and limits in place (QuadraticCost.increase_QF / shift_QF_start, TrajoptCost.py:85-104;
    set_*_limits on an existing TrajoptConstraint).  A cached device workspace uploads those once, so the cache is keyed on their
    content: after a mutation the next solve must equal a FRESH solver's, not the stale one's."""
    N = 12
    (plant, pc, _), _ = make_pair("arm3", N, oracle_models)
    solver = t.TrajoptMPCReference(plant, pc)
    opts = {"expected_reduction_min_SQP_DDP": -100}
    x0, u0 = np.zeros((6, N)), np.zeros((3, N - 1))
    xa = solver.SQP(x0, u0, N, 0.1, t.SQPSolverMethods.PCG_SS, dict(opts))[0]
    pc.increase_QF(4.0)
    xb = solver.SQP(x0, u0, N, 0.1, t.SQPSolverMethods.PCG_SS, dict(opts))[0]
    fresh_cost = t.QuadraticCost(pc.Q.copy(), pc.QF.copy(), pc.R.copy(), pc.xg.copy())
    xf = t.TrajoptMPCReference(plant, fresh_cost).SQP(x0, u0, N, 0.1, t.SQPSolverMethods.PCG_SS, dict(opts))[0]
    assert np.array_equal(xb, xf) and not np.array_equal(xa, xb)
    # the per-knot cost callback follows the mutated weights too
    xk = np.linspace(-0.3, 0.4, 6)
    assert abs(pc.value(xk) - fresh_cost.value(xk)) < 1e-15
    # limits added AFTER the first solve are honoured
    cons = t.TrajoptConstraint(3, 3, 3, N)
    solver2 = t.TrajoptMPCReference(plant, fresh_cost, cons)
    r_free = solver2.SQP(x0, u0, N, 0.1, t.SQPSolverMethods.PCG_SS, dict(opts))
    cons.set_torque_limits([0.2], [-0.2], "QUADRATIC_PENALTY", {})
    r_lim = solver2.SQP(x0, u0, N, 0.1, t.SQPSolverMethods.PCG_SS, dict(opts))
    assert np.max(np.abs(r_free[1])) > 0.25 and np.max(np.abs(r_lim[1])) < np.max(np.abs(r_free[1]))
    cons.set_torque_limits([0.1], [-0.1], "QUADRATIC_PENALTY", {})
    r_lim2 = solver2.SQP(x0, u0, N, 0.1, t.SQPSolverMethods.PCG_SS, dict(opts))
    assert not np.array_equal(r_lim[1], r_lim2[1])


@pytest.mark.parametrize("name,N,B,iters", [("arm3", 12, 6, 10), ("arm2", 10, 4, 8)])
def test_end_effector_cost_exact_hessian_mode(name, N, B, iters, oracle_models):
    """UrdfCost.hess_mode = 1 (TrajoptCost.py:391-395, 494-499): the reference's branch crashes (`hess_x` unbound, no d2Jdq2), so the
    exact Hessian  J_tot^T Q J_tot + sum_i (Q e)_i Hessian(e_i)  is the specification (oracle/cost.py second_order_term; finite-
    difference checked in tests/test_hostemu.py).  Callback against the oracle, then complete solves with the exact linear solver."""
    import copy
    m = oracle_models[name]
    n = m.n
    rng = np.random.default_rng(5)
    ang = rng.uniform(0.4, 2.6, B); rad = rng.uniform(0.6, 0.8 * n, B)
    xg = np.stack([rad * np.cos(ang), rad * np.sin(ang), np.zeros(B), np.zeros(B)], axis=1)
    (plant, pc, _), (_, oc, _) = make_pair(name, N, oracle_models, xg=xg[0], cost_kind="urdf")
    xk = rng.uniform(-1, 1, 2 * n); uk = rng.uniform(-1, 1, n)
    X = np.stack([xk, xk]); U = uk[None]
    H0 = pc.hessian(xk, uk)
    pc.hess_mode = 1; oc.hess_mode = 1
    H1 = pc.hessian(xk, uk)
    assert np.allclose(H1, oc.hessians(X, U)[0], rtol=1e-12, atol=1e-12)
    assert np.max(np.abs(H1 - H0)) > 1e-3 * np.max(np.abs(H0))        # the mode switch reaches the device (cached probe re-keyed on content)
    assert np.allclose(pc.hessian(xk), oc.hessians(X, U)[1][:2 * n, :2 * n], rtol=1e-12, atol=1e-12)
    solver = t.TrajoptMPCReference(plant, pc)
    opts = {"expected_reduction_min_SQP_DDP": -100, "max_iter_SQP_DDP": iters}
    r = solver.solve_batch(np.zeros((B, 2 * n, N)), np.zeros((B, n, N - 1)), xg, N, 0.1, t.SQPSolverMethods.S, dict(opts))
    same = 0
    for b in range(B):
        oc_b = copy.copy(oc); oc_b.xg = xg[b]
        ro = sqp.sqp(m, oc_b, None, np.zeros((2 * n, N)), np.zeros((n, N - 1)), N, 0.1, "S", dict(opts))
        ok = (ro["exit_sqp"], ro["sqp_iter"]) == (r.exit_sqp[b], r.sqp_iter[b]) and sum(ro["ls_trials"]) == r.total_trials[b]
        same += int(ok)
        print("%s hess_mode 1 instance %d: oracle (exit, iters, trials) %s gpu %s rel J %.1e" %
              (name, b, (ro["exit_sqp"], ro["sqp_iter"], sum(ro["ls_trials"])), (int(r.exit_sqp[b]), int(r.sqp_iter[b]), int(r.total_trials[b])),
               abs(ro["J"] - r.J[b]) / max(1.0, abs(ro["J"]))))
        if ok:
            assert abs(ro["J"] - r.J[b]) < 1e-6 * max(1.0, abs(ro["J"]))
    assert same >= B - 1


def test_more_than_103_sqp_iterations_allowed(oracle_models):
    """max_iter_SQP_DDP above the 104 trace rows kept per outer iteration: the solve runs (round 1 refused it), counts and results
    follow the oracle; `trace` keeps the first 104 rows."""
    N = 20
    limits = {"torque": ([0.1], [-0.1], "QUADRATIC_PENALTY")}
    (plant, pc, pcons), (m, oc, ocn) = make_pair("pend", N, oracle_models, limits=limits)
    opts = {"expected_reduction_min_SQP_DDP": -100, "max_iter_SQP_DDP": 150, "max_iter_softConstraints": 3}
    solver = t.TrajoptMPCReference(plant, pc, pcons)
    x, u, e1, e2, outer, it = solver.SQP(np.zeros((2, N)), np.zeros((1, N - 1)), N, 0.1, t.SQPSolverMethods.PCG_SS, options=dict(opts))
    ro = sqp.sqp(m, oc, ocn, np.zeros((2, N)), np.zeros((1, N - 1)), N, 0.1, "PCG-SS", dict(opts))
    assert (e1, e2, outer, it) == (ro["exit_sqp"], ro["exit_soft"], ro["outer_iter"], ro["sqp_iter"])
    assert int(solver.last_result.total_qp[0]) == len(ro["pcg_iters"])
    assert np.max(np.abs(x - ro["x"])) < 1e-9 * max(1.0, np.max(np.abs(ro["x"]))) and len(solver.trace) <= 104
