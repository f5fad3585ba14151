"""Throughput of the iLQR path on BASELINE config 1 (cart-pole, AL torque / velocity limits, batch 1024) and on arm6."""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, "/root/repo")
import trajoptmpcreference_b200 as t


def run(name, N, dt, B, mk_cost, mk_cons, xg, u0, steps=3):
    plant = t.URDFPlant(options={"path_to_urdf": name})
    n = plant.get_num_pos()
    cost = mk_cost(n)
    cons = mk_cons(n, N)
    s = t.BatchSolver(plant, cost, cons, N=N, dt=dt, batch=B)
    x0 = np.zeros((B, 2 * n, N)); U0 = np.broadcast_to(u0.reshape(1, n, 1), (B, n, N - 1)).copy()
    times = []
    s.set_profiling(True)
    for it in range(steps + 2):
        if cons is not None:
            s.reset_multipliers()
        s.set_goals(xg); s.set_trajectory(x0, U0)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        s.solve_ilqr({"max_iter_softConstraints": 6})
        torch.cuda.synchronize(); el = time.perf_counter() - t0
        if it >= 2:
            times.append(el)
    r = s.result()
    print(json.dumps({"workload": "iLQR %s N=%d batch %d" % (name, N, B), "solves_per_s": B / np.mean(times), "ms_per_step": 1e3 * np.mean(times),
                      "iters_per_instance": float(r.total_qp.mean()), "trials_per_instance": float(r.total_trials.mean()),
                      "exit_sqp_hist": np.bincount(r.exit_sqp, minlength=4).tolist(), "exit_soft_hist": np.bincount(r.exit_soft, minlength=4).tolist(),
                      "kernel_ms": {k: round(v[0] * 1e3, 2) for k, v in s.kernel_times().items()}}))


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    B = 1024
    xg = np.zeros((B, 4)); xg[:, 0] = rng.uniform(-0.5, 0.5, B); xg[:, 1] = np.pi

    def cp_cons(n, N):
        c = t.TrajoptConstraint(n, n, n, N)
        c.set_torque_limits([12.0, 1.0], [-12.0, -1.0], "AUGMENTED_LAGRANGIAN", {})
        c.set_velocity_limits([4.0, 8.0], [-4.0, -8.0], "AUGMENTED_LAGRANGIAN", {})
        return c
    run("cartpole", 40, 0.05, B, lambda n: t.QuadraticCost(np.diag([1, 1, 0.1, 0.1]), np.diag([100, 100, 10, 10.0]), np.diag([0.01, 10.0]), np.zeros(4)),
        cp_cons, xg, np.array([0.01, 0.01]))
    B = 2048
    xg6 = np.zeros((B, 12)); xg6[:, :6] = rng.uniform(-0.5, 0.5, (B, 6))
    run("arm6", 64, 0.1, B, lambda n: t.QuadraticCost(np.eye(12), 100.0 * np.eye(12), 0.1 * np.eye(6), np.zeros(12)), lambda n, N: None, xg6, np.zeros(6))
