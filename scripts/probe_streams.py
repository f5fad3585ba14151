"""Does splitting the batch over k solver handles on k CUDA streams (driven by k host threads) shorten the step?  The PCG kernel
holds one instance per SM, so a launch with fewer active instances than SMs leaves SMs idle; concurrent sub-batches fill them."""
import sys, os, time, threading, ctypes
import numpy as np
import torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import trajoptmpcreference_b200 as t
from trajoptmpcreference_b200 import _lib

N, B = 64, int(sys.argv[1]) if len(sys.argv) > 1 else 8192
limits = int(sys.argv[2]) if len(sys.argv) > 2 else 1
plant = t.URDFPlant(options={"path_to_urdf": "arm6"})
cost = t.QuadraticCost(np.eye(12), 100.0 * np.eye(12), 0.1 * np.eye(6), np.zeros(12))
def mkcons():
    if not limits:
        return None
    c = t.TrajoptConstraint(6, 6, 6, N)
    c.set_torque_limits([1.0], [-1.0], "QUADRATIC_PENALTY", {})
    c.set_joint_limits([0.45], [-0.45], "QUADRATIC_PENALTY", {})
    return c
rng = np.random.default_rng(1)
xg = np.zeros((B, 12)); xg[:, :6] = rng.uniform(-0.5, 0.5, (B, 6))
opts = {"expected_reduction_min_SQP_DDP": -100}
for k in (1, 2, 4):
    Bk = B // k
    solvers, streams = [], []
    for i in range(k):
        s = t.BatchSolver(plant, cost, mkcons(), N=N, dt=0.1, batch=Bk)
        st = torch.cuda.Stream()
        _lib.check(s.lib, s.lib.b2t_set_stream(s._h, ctypes.c_void_p(st.cuda_stream)))
        solvers.append(s); streams.append(st)
    def prep():
        for i, s in enumerate(solvers):
            if limits: s.reset_multipliers()
            s.set_goals(xg[i * Bk:(i + 1) * Bk]); s.set_trajectory(np.zeros((Bk, 12, N)), np.zeros((Bk, 6, N - 1)))
    times = []
    for rep in range(3):
        prep(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        th = [threading.Thread(target=s.solve, args=(t.SQPSolverMethods.PCG_SS, dict(opts))) for s in solvers]
        [x.start() for x in th]; [x.join() for x in th]
        torch.cuda.synchronize()
        times.append(time.perf_counter() - t0)
    J = np.concatenate([s.result().J for s in solvers])
    print("streams", k, "ms/step", round(1e3 * min(times[1:]), 1), "solves/s", round(B / min(times[1:])), "sum J", repr(float(J.sum())))
    for s in solvers: s.close()
