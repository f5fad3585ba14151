"""Small end-to-end cases for compute-sanitizer (memcheck / racecheck): every kernel family, both KKT paths, limits, N/S methods."""
import sys
import numpy as np
sys.path.insert(0, "/root/repo")
import trajoptmpcreference_b200 as t

def run(name, N, B, limits, method, dense=False):
    plant = t.URDFPlant(options={"path_to_urdf": name})
    n = plant.get_num_pos()
    cost = t.QuadraticCost(np.eye(2 * n), 100.0 * np.eye(2 * n), 0.1 * np.eye(n), np.zeros(2 * n))
    cons = None
    if limits:
        cons = t.TrajoptConstraint(n, n, n, N)
        cons.set_torque_limits([0.5], [-0.5], "AUGMENTED_LAGRANGIAN", {})
        cons.set_joint_limits([0.45], [-0.45], "QUADRATIC_PENALTY", {})
    s = t.BatchSolver(plant, cost, cons, N=N, dt=0.1, batch=B, dense_kkt=dense)
    rng = np.random.default_rng(0)
    xg = np.zeros((B, 2 * n)); xg[:, :n] = rng.uniform(-0.5, 0.5, (B, n))
    s.set_goals(xg); s.set_trajectory(np.zeros((B, 2 * n, N)), np.zeros((B, n, N - 1)))
    s.solve(method, {"expected_reduction_min_SQP_DDP": -100, "max_iter_softConstraints": 3})
    r = s.result()
    print(name, N, B, limits, method, dense, "exit", np.bincount(r.exit_sqp, minlength=4).tolist(), "J", float(r.J.mean()))

run("arm6", 12, 5, True, t.SQPSolverMethods.PCG_SS)
run("arm6", 12, 5, False, t.SQPSolverMethods.PCG_BJ, dense=True)
run("arm2", 7, 3, True, t.SQPSolverMethods.S)
run("arm3", 9, 4, True, t.SQPSolverMethods.PCG_J)
