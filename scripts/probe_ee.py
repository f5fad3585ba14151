"""Divergence growth between the GPU path and the oracle on the n-link end-effector cost (arm6, method S): J and max|dx| after
k = 1..12 SQP iterations.  Growth by a constant factor per iteration = amplification of rounding, not a defect."""
import copy, sys, os
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests")); sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import trajoptmpcreference_b200 as t
from oracle import rbd, sqp, cost as ocost
from trajoptmpcreference_b200 import model as M
name, N, B = "arm6", 16, 4
m = rbd.Model(M.extract_model(M.builtin_urdf(name))); n = m.n
rng = np.random.default_rng(21)
ang = rng.uniform(0.3, 2.8, B); rad = rng.uniform(0.5, 0.9 * n, B)
xg = np.stack([rad * np.cos(ang), rad * np.sin(ang), np.zeros(B), np.zeros(B)], axis=1)
plant = t.URDFPlant(options={"path_to_urdf": name})
pc = t.UrdfCost(plant, np.eye(4), 100 * np.eye(4), 0.1 * np.eye(n), xg[0])
oc = ocost.UrdfCost(m, np.eye(4), 100 * np.eye(4), 0.1 * np.eye(n), xg[0])
solver = t.TrajoptMPCReference(plant, pc)
for k in (1, 2, 3, 4, 6, 8, 12):
    opts = {"expected_reduction_min_SQP_DDP": -100, "max_iter_SQP_DDP": k}
    r = solver.solve_batch(np.zeros((B, 2 * n, N)), np.zeros((B, n, N - 1)), xg, N, 0.1, t.SQPSolverMethods.S, dict(opts))
    out = []
    for b in range(B):
        oc_b = copy.copy(oc); oc_b.xg = xg[b]
        ro = sqp.sqp(m, oc_b, None, np.zeros((2 * n, N)), np.zeros((n, N - 1)), N, 0.1, "S", dict(opts))
        out.append("%d/%d dJ=%.1e dx=%.1e" % (sum(ro["ls_trials"]), r.total_trials[b], abs(ro["J"] - r.J[b]) / abs(ro["J"]), np.max(np.abs(ro["x"] - r.x[b]))))
    print(k, " | ".join(out))
