import sys, json, numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import trajoptmpcreference_b200 as t
plant = t.URDFPlant(options={"path_to_urdf": "arm6"})
n=6; N=16; B=8
pc = t.QuadraticCost(np.eye(12), 100*np.eye(12), 0.1*np.eye(6), np.zeros(12))
rng = np.random.default_rng(5); xg = np.zeros((B,12)); xg[:, :6] = rng.uniform(-0.5,0.5,(B,6))
s = t.TrajoptMPCReference(plant, pc)
o = {"expected_reduction_min_SQP_DDP": -100}
for dt_ in ("f64","f32"):
    r = s.solve_batch(np.zeros((B,12,N)), np.zeros((B,6,N-1)), xg, N, 0.1, t.SQPSolverMethods.PCG_SS, dict(o), dtype=dt_)
    print(dt_, "J", np.round(r.J,6), "c", np.round(r.c,5), "sqp", r.sqp_iter, "pcg", r.total_pcg, "trials", r.total_trials, "exit", r.exit_sqp)
for tol in (1e-4, 1e-3):
    oo = dict(o); oo["exit_tolerance_linSys"] = tol
    r = s.solve_batch(np.zeros((B,12,N)), np.zeros((B,6,N-1)), xg, N, 0.1, t.SQPSolverMethods.PCG_SS, oo, dtype="f32")
    print("f32 tol", tol, "J", np.round(r.J,6), "sqp", r.sqp_iter, "pcg", r.total_pcg, "exit", r.exit_sqp)
