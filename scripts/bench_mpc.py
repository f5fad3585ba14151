#!/usr/bin/env python
"""Throughput of the batched receding-horizon MPC sweep (BASELINE.json configs[4] wording: "arm6.urdf batched MPC sweep"; SURVEY.md
8f-1): arm6, N=64, the bench workload's goals and box limits, `--batch` closed loops, `--steps` MPC steps each (solve from the warm
start -> apply u_0 -> simulate -> shift trajectories and multipliers on the device).  The first MPC step is the cold solve of
bench.py; every later step starts from the shifted previous solution.  Prints one JSON line with MPC solves/s overall and for the
warm-started steps alone."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench                                   # noqa: E402
import trajoptmpcreference_b200 as t            # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8192)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--limits", type=int, default=1)
    a = ap.parse_args()
    import torch
    N = bench.N_KNOTS
    plant = t.URDFPlant(options={"path_to_urdf": "arm6"})
    cost = t.QuadraticCost(np.eye(12), 100.0 * np.eye(12), 0.1 * np.eye(6), np.zeros(12))
    cons = None
    if a.limits:
        cons = t.TrajoptConstraint(6, 6, 6, N)
        cons.set_torque_limits([bench.U_LIM], [-bench.U_LIM], "QUADRATIC_PENALTY", {})
        cons.set_joint_limits([bench.Q_LIM], [-bench.Q_LIM], "QUADRATIC_PENALTY", {})
    solver = t.TrajoptMPCReference(plant, cost, cons) if cons is not None else t.TrajoptMPCReference(plant, cost)
    xg = bench.workload_goals(1, 0, a.batch)
    x_start = np.zeros((a.batch, 12))
    def timed(steps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = solver.mpc_batch(x_start, xg, N, bench.DT, steps, t.SQPSolverMethods.PCG_SS, dict(bench.SOLVER_OPTS))
        torch.cuda.synchronize()
        return time.perf_counter() - t0, r

    timed(1)                                   # warm-up (allocations, module load)
    cold, _ = timed(1)                         # one MPC step = the cold solve of bench.py + shift
    total, r = timed(a.steps)
    warm = total - cold
    out = {"workload": "MPC sweep: arm6 N=64 %s, %d closed loops x %d MPC steps, PCG-SS" % ("penalty box limits" if a.limits else "no limits", a.batch, a.steps),
           "mpc_solves_per_s": a.batch * a.steps / total, "seconds": total, "cold_first_step_s": cold,
           "warm_mpc_solves_per_s": a.batch * (a.steps - 1) / warm if a.steps > 1 else None, "warm_step_s_mean": warm / max(1, a.steps - 1),
           "sqp_iterations_per_step_mean": [float(v) for v in np.asarray(r.sqp_iter).mean(axis=0)],
           "dist_to_goal_q_mean_first_last": [float(v) for v in np.abs(np.asarray(r.x_closed)[:, :6, :] - xg[:, :6, None]).mean(axis=(0, 1))[[0, -1]]]}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
