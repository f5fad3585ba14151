"""-m gpu: receding-horizon MPC loop (device-resident warm start, b2t_mpc_shift) against oracle/mpc.py (repository spec)."""
import copy

import numpy as np
import pytest

import trajoptmpcreference_b200 as t
from gpu_common import make_pair
from oracle import mpc as ompc

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,N,limits,integ", [("arm2", 10, None, 0), ("pend", 12, {"torque": ([0.3], [-0.3], "AUGMENTED_LAGRANGIAN")}, 0),
                                                 ("arm2", 10, None, 2), ("pend", 12, {"torque": ([0.3], [-0.3], "QUADRATIC_PENALTY")}, 3)])
def test_mpc_loop_vs_oracle(name, N, limits, integ, oracle_models):
    """integ 2 / 3: the plant simulation between solves (k_mpc_shift) and the solves themselves use the reference's midpoint / rk3."""
    steps, B = 5, 3
    (plant, pc, pcons), (m, oc, ocn) = make_pair(name, N, oracle_models, limits=limits, cost_kind="quadratic", integrator=integ,
                                                 xg=np.array([0.4, -0.3, 0, 0]) if name == "arm2" else None)
    n = m.n
    rng = np.random.default_rng(2)
    xg = np.tile(np.asarray(pc.xg, dtype=float), (B, 1)); xg[1:, :n] += rng.uniform(-0.2, 0.2, (B - 1, n))
    xs = np.zeros((B, 2 * n)); xs[:, :n] = rng.uniform(-0.1, 0.1, (B, n))
    solver = t.TrajoptMPCReference(plant, pc, pcons) if pcons is not None else t.TrajoptMPCReference(plant, pc)
    opts = {"expected_reduction_min_SQP_DDP": -100, "max_iter_softConstraints": 3}
    r = solver.mpc_batch(xs, xg, N, 0.1, steps, t.SQPSolverMethods.PCG_SS, dict(opts))
    for b in range(B):
        ocb = copy.copy(oc); ocb.xg = xg[b]
        ro = ompc.mpc(m, ocb, copy.deepcopy(ocn), xs[b], N, 0.1, steps, "PCG-SS", dict(opts), integrator_type=integ)
        assert ro["sqp_iter"] == r.sqp_iter[b].tolist()
        assert np.max(np.abs(ro["x_closed"] - r.x_closed[b])) < 1e-5
        assert np.max(np.abs(ro["u_applied"] - r.u_applied[b])) < 1e-4
