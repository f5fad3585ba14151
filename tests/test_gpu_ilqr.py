"""-m gpu: iLQR kernels (k_ilqr_cost / k_ilqr_backward / k_ilqr_search) against the iLQR oracle (the repository's own
specification; the reference has no iLQR code)."""
import copy

import numpy as np
import pytest

import trajoptmpcreference_b200 as t
from oracle import rbd, cost as ocost, constraint as ocons, ilqr
from trajoptmpcreference_b200.model import extract_model, builtin_urdf

pytestmark = pytest.mark.gpu


def _cartpole(N, limits):
    m = rbd.Model(extract_model(builtin_urdf("cartpole")))
    Q, QF, R = np.diag([1, 1, 0.1, 0.1]), np.diag([100, 100, 10, 10.0]), np.diag([0.01, 10.0])
    xg = np.array([0, np.pi, 0, 0])
    plant = t.URDFPlant(options={"path_to_urdf": "cartpole"})
    pc = t.QuadraticCost(Q.copy(), QF.copy(), R.copy(), xg.copy())
    oc = ocost.QuadraticCost(Q, QF, R, xg)
    pcons = ocn = None
    if limits:
        pcons = t.TrajoptConstraint(2, 2, 2, N); ocn = ocons.SoftConstraints(2, 2, 2, N)
        for c in (pcons, ocn):
            c.set_torque_limits([12.0, 1.0], [-12.0, -1.0], "AUGMENTED_LAGRANGIAN")
            c.set_velocity_limits([4.0, 8.0], [-4.0, -8.0], "AUGMENTED_LAGRANGIAN")
    return m, plant, pc, oc, pcons, ocn


@pytest.mark.parametrize("limits", [False, True])
def test_cartpole_ilqr_vs_oracle(limits):
    """BASELINE config 1 (cart-pole iLQR with augmented-Lagrangian torque / velocity limits) at batch 1."""
    N, dt = 40, 0.05
    m, plant, pc, oc, pcons, ocn = _cartpole(N, limits)
    solver = t.TrajoptMPCReference(plant, pc, pcons) if pcons is not None else t.TrajoptMPCReference(plant, pc)
    opts = {"max_iter_softConstraints": 6}
    x, u, e1, e2, outer, it = solver.iLQR(np.zeros((4, N)), 0.01 * np.ones((2, N - 1)), N, dt, dict(opts))
    ro = ilqr.ilqr(m, oc, ocn, np.zeros((4, N)), 0.01 * np.ones((2, N - 1)), N, dt, dict(opts))
    assert (e1, e2, outer, it) == (ro["exit_sqp"], ro["exit_soft"], ro["outer_iter"], ro["sqp_iter"])
    r = solver.last_result
    assert (int(r.total_qp[0]), int(r.total_trials[0])) == (ro["total_iters"], ro["total_trials"])
    assert abs(r.J[0] - ro["J"]) < 1e-7 * abs(ro["J"])
    assert np.max(np.abs(x - ro["x"])) < 1e-5 and np.max(np.abs(u - ro["u"])) < 1e-4
    assert abs(x[1, -1] - np.pi) < 0.05


@pytest.mark.parametrize("integ", [0, 2, 3])
def test_ilqr_batch_arm6_vs_oracle(integ, oracle_models):
    """integ 2 / 3: rollouts through the reference's midpoint / rk3 step, Riccati pass on the stored full [A B] (k_ab_multi)."""
    N, B = 16, 12
    m = oracle_models["arm6"]
    Q, QF, R = np.eye(12), 100.0 * np.eye(12), 0.1 * np.eye(6)
    plant = t.URDFPlant(integrator_type=integ, options={"path_to_urdf": "arm6"})
    pc = t.QuadraticCost(Q.copy(), QF.copy(), R.copy(), np.zeros(12))
    pcons = t.TrajoptConstraint(6, 6, 6, N); ocn = ocons.SoftConstraints(6, 6, 6, N)
    for c in (pcons, ocn):
        c.set_torque_limits([1.0], [-1.0], "QUADRATIC_PENALTY")
    rng = np.random.default_rng(3)
    xg = np.zeros((B, 12)); xg[:, :6] = rng.uniform(-0.5, 0.5, (B, 6))
    solver = t.TrajoptMPCReference(plant, pc, pcons)
    opts = {"max_iter_softConstraints": 3}
    r = solver.ilqr_batch(np.zeros((B, 12, N)), np.zeros((B, 6, N - 1)), xg, N, 0.1, dict(opts))
    same = 0
    for b in range(B):
        ro = ilqr.ilqr(m, ocost.QuadraticCost(Q, QF, R, xg[b]), copy.deepcopy(ocn), np.zeros((12, N)), np.zeros((6, N - 1)), N, 0.1, dict(opts),
                       integrator_type=integ)
        ok = (ro["exit_sqp"], ro["exit_soft"], ro["outer_iter"], ro["sqp_iter"], ro["total_iters"], ro["total_trials"]) == \
             (r.exit_sqp[b], r.exit_soft[b], r.outer_iter[b], r.sqp_iter[b], r.total_qp[b], r.total_trials[b])
        same += int(ok)
        if ok:
            assert abs(ro["J"] - r.J[b]) < 1e-7 * max(1.0, abs(ro["J"]))
            assert np.max(np.abs(ro["x"] - r.x[b])) < 1e-5
    assert same >= int(0.9 * B)
