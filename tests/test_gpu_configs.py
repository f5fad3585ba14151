"""-m gpu: the BASELINE.json configurations at their full sizes (size-independent properties + oracle spot checks).
  [0] pendulum SQP-PCG, single instance                -> tests/test_gpu_solve.py (pend_* golden cases)
  [1] cart-pole iLQR, AL torque/velocity limits, batch 1024
  [2] arm2 SQP-PCG reaching task (UrdfCost), N=32, batch 4096
  [3] arm6 SQP-PCG, penalty box limits, N=64, batch 8192  (bench.py default; reduced batch in test_gpu_solve.py)
"""
import copy

import numpy as np
import pytest

import trajoptmpcreference_b200 as t
from gpu_common import make_pair
from oracle import sqp, ilqr, rbd, cost as ocost, constraint as ocons
from trajoptmpcreference_b200.model import extract_model, builtin_urdf

pytestmark = pytest.mark.gpu


def test_config2_arm2_urdfcost_N32_batch4096(oracle_models):
    """goals on the reachable disc, r ~ U(0.5,1.9), theta ~ U(0,2pi), default_rng(0)."""
    N, B = 32, 4096
    rng = np.random.default_rng(0)
    r_, th = rng.uniform(0.5, 1.9, B), rng.uniform(0, 2 * np.pi, B)
    xg = np.stack([r_ * np.cos(th), r_ * np.sin(th), np.zeros(B), np.zeros(B)], axis=1)
    (plant, pc, _), (m, oc, _) = make_pair("arm2", N, oracle_models)
    solver = t.TrajoptMPCReference(plant, pc)
    opts = {"expected_reduction_min_SQP_DDP": -100}
    x0 = np.zeros((B, 4, N)); u0 = np.zeros((B, 2, N - 1))
    r = solver.solve_batch(x0, u0, xg, N, 0.1, t.SQPSolverMethods.PCG_SS, dict(opts))
    assert set(np.unique(r.exit_sqp)) <= {1, 2, 3} and np.all(np.isfinite(r.x)) and np.all(np.isfinite(r.J))
    # spot checks against the oracle
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    from make_c4_fixture import _perturber          # S -> S (1 +/- 2^-52): the oracle's own sensitivity sets the tolerance
    same = 0
    idx = [0, 1, 7, 100, 2047, 4095]
    for b in idx:
        ocb = copy.copy(oc); ocb.xg = xg[b]
        ro = sqp.sqp(m, ocb, None, np.zeros((4, N)), np.zeros((2, N - 1)), N, 0.1, "PCG-SS", dict(opts))
        rp = sqp.sqp(m, ocb, None, np.zeros((4, N)), np.zeros((2, N - 1)), N, 0.1, "PCG-SS", dict(opts, _perturb_S=_perturber(b)))
        stable = rp["pcg_iters"] == ro["pcg_iters"] and rp["ls_trials"] == ro["ls_trials"]
        fJ = abs(rp["J"] - ro["J"]) / max(1.0, abs(ro["J"])); fx = float(np.max(np.abs(rp["x"] - ro["x"])))
        ok = (ro["exit_sqp"], ro["sqp_iter"], sum(ro["pcg_iters"]), sum(ro["ls_trials"])) == (r.exit_sqp[b], r.sqp_iter[b], r.total_pcg[b], r.total_trials[b])
        same += int(ok)
        eJ = abs(ro["J"] - r.J[b]) / max(1.0, abs(ro["J"])); ex = float(np.max(np.abs(ro["x"] - r.x[b])))
        print("C3 instance %d: counts %s, rel J %.1e (1-ulp floor %.1e), |dx| %.1e (floor %.1e)%s" %
              (b, "same" if ok else "DIFFER", eJ, fJ, ex, fx, "" if stable else "  [oracle counts unstable under 1 ulp]"))
        if ok and stable:
            assert eJ <= max(100 * fJ, 1e-9) and ex <= max(100 * fx, 1e-9), (b, eJ, fJ, ex, fx)
    assert same >= len(idx) - 1
    # permutation invariance at full size
    perm = rng.permutation(B)
    r2 = solver.solve_batch(x0, u0, xg[perm], N, 0.1, t.SQPSolverMethods.PCG_SS, dict(opts))
    assert np.array_equal(np.array(r.x)[perm], r2.x)


def test_config1_cartpole_ilqr_batch1024():
    N, dt, B = 40, 0.05, 1024
    m = rbd.Model(extract_model(builtin_urdf("cartpole")))
    Q, QF, R = np.diag([1, 1, 0.1, 0.1]), np.diag([100, 100, 10, 10.0]), np.diag([0.01, 10.0])
    plant = t.URDFPlant(options={"path_to_urdf": "cartpole"})
    pc = t.QuadraticCost(Q.copy(), QF.copy(), R.copy(), np.zeros(4))
    pcons = t.TrajoptConstraint(2, 2, 2, N); ocn = ocons.SoftConstraints(2, 2, 2, N)
    for c in (pcons, ocn):
        c.set_torque_limits([12.0, 1.0], [-12.0, -1.0], "AUGMENTED_LAGRANGIAN")
        c.set_velocity_limits([4.0, 8.0], [-4.0, -8.0], "AUGMENTED_LAGRANGIAN")
    rng = np.random.default_rng(0)
    xg = np.zeros((B, 4)); xg[:, 0] = rng.uniform(-0.5, 0.5, B); xg[:, 1] = np.pi
    solver = t.TrajoptMPCReference(plant, pc, pcons)
    opts = {"max_iter_softConstraints": 6}
    x0 = np.zeros((B, 4, N)); u0 = 0.01 * np.ones((B, 2, N - 1))
    r = solver.ilqr_batch(x0, u0, xg, N, dt, dict(opts))
    assert np.all(np.isfinite(r.x)) and set(np.unique(r.exit_sqp)) <= {1, 2, 3}
    assert np.mean(np.abs(r.x[:, 1, -1] - np.pi) < 0.1) > 0.95               # the pole is swung up
    assert np.max(np.abs(r.u[:, 0, :])) < 12.3                                  # cart force inside its (soft) limit
    # returned trajectories are rollouts: x_{k+1} = integrator(x_k, u_k)
    matched, spots = 0, (0, 255, 511, 767, 1023)
    for b in spots:
        X = ilqr.rollout(m, r.x[b][:, 0], r.u[b].T, dt)
        assert np.max(np.abs(X.T - r.x[b])) < 1e-8
        ro = ilqr.ilqr(m, ocost.QuadraticCost(Q, QF, R, xg[b]), copy.deepcopy(ocn), np.zeros((4, N)), 0.01 * np.ones((2, N - 1)), N, dt, dict(opts))
        same = (ro["total_iters"], ro["total_trials"]) == (r.total_qp[b], r.total_trials[b])
        print("cart-pole iLQR instance %d: oracle (iterations, rollouts) %s gpu %s  rel J %.1e" %
              (b, (ro["total_iters"], ro["total_trials"]), (int(r.total_qp[b]), int(r.total_trials[b])), abs(ro["J"] - r.J[b]) / abs(ro["J"])))
        matched += int(same)
        if same:
            assert abs(ro["J"] - r.J[b]) < 1e-6 * abs(ro["J"])
    # no silent skip: at least 4 of the 5 spot-checked instances must reproduce the oracle's iteration and rollout counts
    assert matched >= len(spots) - 1, "only %d of %d spot checks reproduce the oracle's counts" % (matched, len(spots))
