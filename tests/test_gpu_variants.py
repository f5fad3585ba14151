"""-m gpu: kernel variants that claim bit-identical results must produce them.

k_schur_rows (n threads per block row; 2- and 3-CTA register allocations) against k_schur_diag (one thread per block row,
B2T_SCHUR_V1=1), and the asynchronous active-count read-back against the blocking one (B2T_SYNC_PASSES=1) on complete solves.
The switches are read when a solver is created / a solve starts, so each variant gets its own BatchSolver."""
import os

import numpy as np
import pytest

import trajoptmpcreference_b200 as t
from gpu_common import make_pair

pytestmark = pytest.mark.gpu


class _env:
    def __init__(self, **kv):
        self.kv = kv

    def __enter__(self):
        self.old = {k: os.environ.get(k) for k in self.kv}
        for k, v in self.kv.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v

    def __exit__(self, *a):
        for k, v in self.old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def _problem(name, N, batch, oracle_models, seed, integrator=0):
    m = oracle_models[name]
    n = m.n
    rng = np.random.default_rng(seed)
    xg = np.concatenate([rng.uniform(-0.5, 0.5, n), np.zeros(n)])
    limits = {"torque": ([1.0] * n, [-1.0] * n, "QUADRATIC_PENALTY"), "joint": ([0.45] * n, [-0.45] * n, "QUADRATIC_PENALTY")}
    (plant, pc, pcons), _ = make_pair(name, N, oracle_models, xg=xg, limits=limits, integrator=integrator)
    x = 0.3 * rng.standard_normal((batch, 2 * n, N))
    u = 0.5 * rng.standard_normal((batch, n, N - 1))
    return plant, pc, pcons, x, u


@pytest.mark.parametrize("name,N", [("arm6", 64), ("arm6", 9), ("arm2", 33), ("pend", 5)])
@pytest.mark.parametrize("rho", [1e-3, 4.0])
def test_schur_variants_bit_identical(name, N, rho, oracle_models):
    batch = 3
    out = {}
    for tag, env in (("rows2", dict(B2T_SCHUR_V1=None, B2T_SCHUR_MINB=None)), ("rows3", dict(B2T_SCHUR_V1=None, B2T_SCHUR_MINB="3")),
                     ("diag", dict(B2T_SCHUR_V1="1", B2T_SCHUR_MINB=None))):
        with _env(**env):
            plant, pc, pcons, x, u = _problem(name, N, batch, oracle_models, seed=N)      # fresh objects: constraints carry state
            s = t.BatchSolver(plant, pc, pcons, N=N, dt=0.1, batch=batch)
            s.set_trajectory(x, u)
            s.set_initial_state(x[:, :, 0] + 0.01)
            s.stage_dynamics()
            s.stage_kkt(rho, t.SQPSolverMethods.S)        # exact method: the sub-diagonal blocks are written too
            out[tag] = [s.fetch(k).copy() for k in ("Sd", "So", "gamma")]
            del s
    for tag in ("rows3", "diag"):
        for a, b, k in zip(out["rows2"], out[tag], ("Sd", "So", "gamma")):
            assert np.array_equal(a, b), (tag, k, float(np.max(np.abs(a - b))))


def test_lagged_count_readback_identical(oracle_models):
    """complete solves: the pass loop that reads the active count one pass late returns exactly what the blocking loop returns"""
    N, batch = 16, 37
    res = {}
    for tag, env in (("lagged", dict(B2T_SYNC_PASSES=None)), ("sync", dict(B2T_SYNC_PASSES="1"))):
        with _env(**env):
            plant, pc, pcons, x, u = _problem("arm6", N, batch, oracle_models, seed=5)
            s = t.BatchSolver(plant, pc, pcons, N=N, dt=0.1, batch=batch)
            s.set_trajectory(np.zeros_like(x), np.zeros_like(u))
            s.set_initial_state(np.zeros((batch, x.shape[1])))
            s.set_goals(np.concatenate([np.random.default_rng(9).uniform(-0.5, 0.5, (batch, 6)), np.zeros((batch, 6))], axis=1))
            s.solve(t.SQPSolverMethods.PCG_SS, options={"expected_reduction_min_SQP_DDP": -100})
            res[tag] = s.result()
            del s
    a, b = res["lagged"], res["sync"]
    assert np.array_equal(a.x, b.x) and np.array_equal(a.u, b.u)
    assert np.array_equal(a.sqp_iter, b.sqp_iter) and np.array_equal(a.exit_sqp, b.exit_sqp) and np.array_equal(a.exit_soft, b.exit_soft)


@pytest.mark.parametrize("N", [64, 23])
def test_pcg_kernel_variants(N, oracle_models):
    """The PCG kernels of the structured path on the same problems: k_pcg3 (default, B2T_PCG_VARIANT=3), k_pcg4 (5: the own-block halves
    of the products ahead of the barriers), k_pcg6 (6: six lanes per knot, different summation order) and k_pcg3 with the column
    form of the D^-1 products (7: partial sums + a reduce-scatter over the knot's lanes): identical iteration counts
    on these problems, solutions equal to rounding.  The alternatives are measured dead ends (profiles/README.md); they stay
    selectable for A/B runs and must stay correct."""
    batch = 5
    res = {}
    for variant in ("3", "5", "6", "7"):
        with _env(B2T_PCG_VARIANT=variant):
            plant, pc, pcons, x, u = _problem("arm6", N, batch, oracle_models, seed=N)
            s = t.BatchSolver(plant, pc, pcons, N=N, dt=0.1, batch=batch)
            assert s.pcg_kernel_name() == {"3": "k_pcg3", "5": "k_pcg4", "6": "k_pcg6", "7": "k_pcg3"}[variant]
            s.set_trajectory(x, u)
            s.stage_dynamics()
            s.stage_kkt(1e-3, t.SQPSolverMethods.PCG_SS)
            it = s.stage_pcg(t.SQPSolverMethods.PCG_SS, 1e-6, 100)
            res[variant] = (np.array(it), s.fetch("l").copy())
            s.close()
    for v in ("5", "6", "7"):
        assert np.array_equal(res["3"][0], res[v][0]), v
        assert np.max(np.abs(res["3"][1] - res[v][1])) < 1e-9 * np.max(np.abs(res["3"][1])), v


@pytest.mark.parametrize("name,N,batch,limits", [("arm6", 64, 5, True), ("arm6", 64, 333, True), ("arm6", 23, 40, False), ("arm4", 10, 7, True),
                                                      ("arm6", 128, 9, True), ("arm6", 100, 150, False), ("arm6", 33, 297, True), ("arm6", 65, 149, True), ("arm6", 64, 151, True),
                                                      ("arm4", 48, 150, True)])
def test_pcg_tensor_memory_kernel_bit_identical(name, N, batch, limits, oracle_models):
    """k_pcg_tm (B2T_PCG_VARIANT=8, the default for fp64 on the structured path when more instances are active than there are SMs:
    matrices in tensor memory, two instances per SM, instances drawn from a ticket counter) against k_pcg3 on the same systems:
    the products are streamed in k_pcg3's order of operations, so iteration counts AND solutions are bit-identical.  333 instances
    exercise the work queue (more instances than 2 x 148 halves); B2T_PCG_TM_MIN=1 forces the kernel for small batches.  Horizons
    64 < N <= 128 run one instance per CTA (k_pcg_tm<512>) against k_pcg3's 512-thread instantiation."""
    res = {}
    for variant in ("3", "8"):
        with _env(B2T_PCG_VARIANT=variant, B2T_PCG_TM_MIN="1"):
            # batch 151: the semi-implicit Euler integrator (tau = dt: the run-time instantiation also at N = 64)
            plant, pc, pcons, x, u = _problem(name, N, batch, oracle_models, seed=N + batch, integrator=1 if batch == 151 else 0)
            s = t.BatchSolver(plant, pc, pcons if limits else None, N=N, dt=0.1, batch=batch)
            assert s.pcg_kernel_name() == {"3": "k_pcg3", "8": "k_pcg_tm"}[variant]
            s.set_trajectory(x, u)
            s.stage_dynamics()
            s.stage_kkt(1e-3, t.SQPSolverMethods.PCG_SS)
            for method in (t.SQPSolverMethods.PCG_SS, t.SQPSolverMethods.PCG_BJ):
                it = s.stage_pcg(method, 1e-6, 100)
                res[variant, method] = (np.array(it), s.fetch("l").copy())
            s.close()
    for method in (t.SQPSolverMethods.PCG_SS, t.SQPSolverMethods.PCG_BJ):
        assert np.array_equal(res["3", method][0], res["8", method][0])
        assert res["3", method][0].min() > 0
        assert np.array_equal(res["3", method][1], res["8", method][1])


@pytest.mark.parametrize("N,batch", [(64, 200), (40, 11), (100, 7)])
def test_tensor_memory_solve_bit_identical(N, batch, oracle_models):
    """Complete SQP solves with k_pcg_tm in every pass (B2T_PCG_TM_MIN=1) against the k_pcg3 pipeline (B2T_PCG_VARIANT=3), and with
    the default hand-over between the two (k_pcg_tm in passes with more active instances than SMs, k_pcg3 below): every count, exit
    code, trace row and trajectory identical -- a solve does not depend on which kernel ran which pass, hence not on what else is
    in the batch.  N = 64 runs the N = 64 / Euler instantiation, 40 the run-time one, 100 one instance per CTA."""
    out = {}
    for tag, env in (("pcg3", dict(B2T_PCG_VARIANT="3")), ("tm", dict(B2T_PCG_VARIANT="8", B2T_PCG_TM_MIN="1")),
                     ("tm_default", dict(B2T_PCG_VARIANT=None, B2T_PCG_TM_MIN=None))):
        with _env(**env):
            plant, pc, pcons, x, u = _problem("arm6", N, batch, oracle_models, seed=5 + N)
            rng = np.random.default_rng(N)
            xg = np.zeros((batch, 12)); xg[:, :6] = rng.uniform(-0.5, 0.5, (batch, 6))
            solver = t.TrajoptMPCReference(plant, pc, pcons)
            r = solver.solve_batch(x, u, xg, N, 0.1, t.SQPSolverMethods.PCG_SS, {"expected_reduction_min_SQP_DDP": -100, "max_iter_softConstraints": 2,
                                                                               "max_iter_SQP_DDP": 12})
            s = solver.batch_solver(N, 0.1, batch)
            assert s.pcg_kernel_name() == ("k_pcg3" if tag == "pcg3" else "k_pcg_tm")
            out[tag] = ({k: np.array(v) for k, v in r.items()}, s.get_trace().copy())
    for tag in ("tm", "tm_default"):
        for k in out["pcg3"][0]:
            assert np.array_equal(out["pcg3"][0][k], out[tag][0][k]), (tag, k)
        assert np.array_equal(out["pcg3"][1], out[tag][1]), tag
    assert out["pcg3"][0]["total_qp"].min() >= 2


@pytest.mark.parametrize("name,N,kind", [("arm6", 64, "limits"), ("arm6", 17, "limits"), ("arm2", 10, "urdf"), ("pend", 20, "hard")])
def test_parallel_line_search_bit_identical(name, N, kind, oracle_models):
    """k_linesearch_par (all trials of a search evaluated at once, chosen for passes with few active instances; B2T_LS_PAR sets the
    threshold) against the sequential k_linesearch on complete solves: every count, exit code, trace row and trajectory identical."""
    out = {}
    for tag, thr in (("seq", "0"), ("par", "1000000")):
        with _env(B2T_LS_PAR=thr):
            if kind == "limits":
                plant, pc, pcons, x, u = _problem(name, N, 6, oracle_models, seed=3 + N)
                method = t.SQPSolverMethods.PCG_SS
            elif kind == "urdf":
                (plant, pc, pcons), _ = make_pair(name, N, oracle_models)
                x = np.zeros((6, 4, N)); u = np.zeros((6, 2, N - 1)); method = t.SQPSolverMethods.PCG_SS
            else:
                (plant, pc, pcons), _ = make_pair(name, N, oracle_models, limits={"torque": ([0.5], [-0.5], "ACTIVE_SET")})
                x = np.zeros((6, 2, N)); u = np.zeros((6, 1, N - 1)); method = t.SQPSolverMethods.S
            n = plant.get_num_pos()
            rng = np.random.default_rng(9)
            xg = np.zeros((6, pc.xg.size)); xg[:] = pc.xg; xg[:, :min(n, 2)] += 0.1 * rng.standard_normal((6, min(n, 2)))
            solver = t.TrajoptMPCReference(plant, pc, pcons) if pcons is not None else t.TrajoptMPCReference(plant, pc)
            r = solver.solve_batch(x, u, xg, N, 0.1, method, {"expected_reduction_min_SQP_DDP": -100, "max_iter_softConstraints": 3})
            s = solver.batch_solver(N, 0.1, 6)
            out[tag] = ({k: np.array(v) for k, v in r.items()}, s.get_trace().copy())
    a, b = out["seq"], out["par"]
    for k in a[0]:
        assert np.array_equal(a[0][k], b[0][k]), k
    assert np.array_equal(a[1], b[1])
    assert a[0]["total_trials"].sum() > a[0]["total_qp"].sum()          # searches with more than one trial were exercised
