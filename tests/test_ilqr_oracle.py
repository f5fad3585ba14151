"""iLQR oracle (oracle/ilqr.py -- UNPINNED: the reference has no iLQR code).  Consistency checks the
specification can offer without a reference: agreement with the (pinned) SQP oracle at the optimum of an unconstrained problem,
the backward pass against the KKT solution of the same linearisation, monotone cost."""
import numpy as np

from oracle import rbd, plant, cost as ocost, constraint as ocons, sqp, ilqr, kkt, dense
from trajoptmpcreference_b200.model import extract_model, builtin_urdf


def test_ilqr_and_sqp_reach_the_same_optimum(oracle_models):
    m = oracle_models["arm2"]
    c = ocost.QuadraticCost(np.eye(4), 100 * np.eye(4), 0.1 * np.eye(2), np.array([0.4, -0.3, 0, 0]))
    N = 16
    ri = ilqr.ilqr(m, c, None, np.zeros((4, N)), np.zeros((2, N - 1)), N, 0.1, {"exit_tolerance_SQP_DDP": 1e-10})
    rs = sqp.sqp(m, c, None, np.zeros((4, N)), np.zeros((2, N - 1)), N, 0.1, "N",
                 {"exit_tolerance_SQP_DDP": 1e-10, "expected_reduction_min_SQP_DDP": -100, "max_iter_SQP_DDP": 60})
    assert rs["c"] < 1e-6
    assert abs(ri["J"] - rs["J"]) < 1e-8 * abs(rs["J"])
    assert np.max(np.abs(ri["x"] - rs["x"])) < 1e-7 and np.max(np.abs(ri["u"] - rs["u"])) < 1e-7
    Js = [t["J"] for t in ri["trace"] if t["succeeded_line_search"]]
    assert all(a >= b for a, b in zip(Js, Js[1:]))


def test_backward_pass_equals_kkt_step_on_a_feasible_trajectory(oracle_models):
    """With zero defects the iLQR step (alpha = 1, linear rollout) is the solution of the same QP the SQP path solves."""
    m = oracle_models["arm3"]
    n, nx, N, dt = 3, 6, 8, 0.1
    rng = np.random.default_rng(4)
    U = rng.uniform(-0.002, 0.002, (N - 1, n))      # the light planar arms are violently unstable under explicit Euler
    X = ilqr.rollout(m, rng.uniform(-0.05, 0.05, nx), U, dt)
    assert np.max(np.abs(X)) < 5
    c = ocost.QuadraticCost(np.eye(nx), 100 * np.eye(nx), 0.1 * np.eye(n), np.concatenate([np.linspace(0.5, -0.5, n), np.zeros(n)]))
    A, B = plant.integrator(m, X[:N - 1], U, dt, 0, True)
    g, H = c.gradients(X, U), c.hessians(X, U)
    rho = 1e-3
    kff, K, dV1, dV2, ok = ilqr.backward_pass(A, B, g, H, rho, nx)
    assert ok and dV1 < 0
    # linear closed-loop rollout of the step
    dx = np.zeros((N, nx)); du = np.zeros((N - 1, n))
    for k in range(N - 1):
        du[k] = kff[k] + K[k] @ dx[k]
        dx[k + 1] = A[k] @ dx[k] + B[k] @ du[k]
    # the same QP through the KKT system, regularising only the control blocks (iLQR adds rho to Quu only)
    blocks = kkt.form_blocks(m, c, None, X, U, X[0], dt)
    for k in range(N - 1):
        blocks["G"][k][nx:, nx:] += rho * np.eye(n)
    sol = dense.kkt_solve_dense(blocks, 0.0, nx)[:, 0]
    mdim = nx + n
    dz = sol[:mdim * (N - 1)].reshape(N - 1, mdim)
    assert np.max(np.abs(-dz[:, :nx] - dx[:N - 1])) < 1e-8      # SQP applies x - dz
    assert np.max(np.abs(-dz[:, nx:] - du)) < 1e-8


def test_cartpole_swingup_with_limits():
    m = rbd.Model(extract_model(builtin_urdf("cartpole")))
    N, dt = 40, 0.05
    c = ocost.QuadraticCost(np.diag([1, 1, 0.1, 0.1]), np.diag([100, 100, 10, 10.0]), np.diag([0.01, 10.0]), np.array([0, np.pi, 0, 0]))
    cons = ocons.SoftConstraints(2, 2, 2, N)
    cons.set_torque_limits([12.0, 1.0], [-12.0, -1.0], "AUGMENTED_LAGRANGIAN")
    cons.set_velocity_limits([4.0, 8.0], [-4.0, -8.0], "AUGMENTED_LAGRANGIAN")
    r = ilqr.ilqr(m, c, cons, np.zeros((4, N)), 0.01 * np.ones((2, N - 1)), N, dt, {"max_iter_softConstraints": 6})
    assert abs(r["x"][1, -1] - np.pi) < 0.05 and np.abs(r["u"][0]).max() < 12.05
    # dynamic feasibility of the returned trajectory (iLQR iterates are rollouts)
    X = ilqr.rollout(m, r["x"][:, 0], r["u"].T, dt)
    assert np.max(np.abs(X.T - r["x"])) < 1e-10
