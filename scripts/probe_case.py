import sys, json, copy, numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import trajoptmpcreference_b200 as t
from oracle import rbd, cost as ocost, sqp
g = json.load(open('/root/repo/tests/golden/models.json'))
name, N = sys.argv[1], int(sys.argv[2])
m = rbd.Model(g[name]); n = m.n
xg0 = np.concatenate([np.linspace(0.5, -0.5, n), np.zeros(n)])
B = 3
xg = np.tile(xg0, (B, 1)); xg[1:, :n] *= np.array([[0.5], [-0.7]])
plant = t.URDFPlant(options={"path_to_urdf": name})
pc = t.QuadraticCost(np.eye(2*n), 100*np.eye(2*n), 0.1*np.eye(n), xg0.copy())
solver = t.TrajoptMPCReference(plant, pc)
opts = {"expected_reduction_min_SQP_DDP": -100, "max_iter_SQP_DDP": 30}
for method, om in ((t.SQPSolverMethods.S, "S"), (t.SQPSolverMethods.N, "N"), (t.SQPSolverMethods.PCG_SS, "PCG-SS")):
    r = solver.solve_batch(np.zeros((B, 2*n, N)), np.zeros((B, n, N-1)), xg, N, 0.1, method, dict(opts))
    for b in range(B):
        ro = sqp.sqp(m, ocost.QuadraticCost(np.eye(2*n), 100*np.eye(2*n), 0.1*np.eye(n), xg[b]), None, np.zeros((2*n, N)), np.zeros((n, N-1)), N, 0.1, om, dict(opts))
        print(om, b, "gpu", r.exit_sqp[b], r.sqp_iter[b], r.total_trials[b], "%.10f" % r.J[b], "| oracle", ro["exit_sqp"], ro["sqp_iter"], sum(ro["ls_trials"]), "%.10f" % ro["J"], [(round(t_["alpha"],4), round(t_["J"],6)) for t_ in ro["trace"][1:6]])
