#!/usr/bin/env python
"""Small driver for ncu captures of the PCG kernels: arm6, N=64, `batch` instances (default 296 = 2 per SM), no limits unless
LIMITS=1, two SQP iterations.  The kernel variant follows B2T_PCG_VARIANT (3: k_pcg3, 5: k_pcg4, 6: k_pcg6)."""
import os
import sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench                                   # noqa: E402
import trajoptmpcreference_b200 as t            # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 296
N = bench.N_KNOTS
plant = t.URDFPlant(options={"path_to_urdf": "arm6"})
cost = t.QuadraticCost(np.eye(12), 100.0 * np.eye(12), 0.1 * np.eye(6), np.zeros(12))
cons = None
if os.environ.get("LIMITS") == "1":
    cons = t.TrajoptConstraint(6, 6, 6, N)
    cons.set_torque_limits([bench.U_LIM], [-bench.U_LIM], "QUADRATIC_PENALTY", {})
    cons.set_joint_limits([bench.Q_LIM], [-bench.Q_LIM], "QUADRATIC_PENALTY", {})
solver = t.TrajoptMPCReference(plant, cost, cons) if cons is not None else t.TrajoptMPCReference(plant, cost)
opts = dict(bench.SOLVER_OPTS); opts["max_iter_SQP_DDP"] = int(os.environ.get("SQP_ITERS", "2")); opts["max_iter_softConstraints"] = 1
r = solver.solve_batch(np.zeros((B, 12, N)), np.zeros((B, 6, N - 1)), bench.workload_goals(1, 0, B), N, bench.DT, t.SQPSolverMethods.PCG_SS, opts)
print("pcg iterations per instance", float(np.mean(r.total_pcg)), "qp", float(np.mean(r.total_qp)))
