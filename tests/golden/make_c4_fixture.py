#!/usr/bin/env python
"""Golden fixture for the HEADLINE workload (BASELINE.json configs[3], SURVEY.md 8d "C4"): fixed arm6, N=64, joint-space
QuadraticCost, quadratic-penalty torque AND joint box limits, method PCG-SS, default options (10 outer iterations).

  python tests/golden/make_c4_fixture.py [instances=64]      ->  tests/golden/c4_fixture.npz

The reference itself crashes on multi-coordinate box limits, so the generator is the oracle restatement
(oracle/sqp.py + oracle/constraint.py: element-wise restatement of TrajoptConstraint.py:53-166,295-378 and the outer update
TrajoptMPCReference.py:483-508) -- "restatement-pinned".  The goals are the FIRST `instances` goals of
bench.workload_goals(1, 0, 8192), i.e. exactly the instances bench.py solves at N=1; the GPU test solves them inside a batch of
8192 and bench.py compares its cpu_baseline sample against the same arrays.

Stored per instance: exit codes, outer_iter, sqp_iter, number of QP solves, sum of PCG iterations, sum of line-search trials,
J, c, x, u, and the per-QP PCG iteration / line-search-trial sequences (padded with -1) plus J at the start of every outer
iteration and the QP-solve / PCG-iteration / trial counts of every outer iteration, so that a mismatch can be located (first outer iteration / QP solve at which the GPU path leaves the oracle's path).
Also stored: the oracle's OWN parity floor on each instance (`floor_*`): the solve repeated with S perturbed by 1 ulp in every QP
solve (FLOOR_SEEDS sign patterns) -- whether the counts survive, and how far J, x, u move.  An instance whose counts do not
survive a 1-ulp perturbation cannot be expected to reproduce them on different hardware arithmetic either.
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
os.environ.setdefault("OMP_NUM_THREADS", "1")
os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
os.environ.setdefault("MKL_NUM_THREADS", "1")

QP_CAP = 1100          # 10 outer iterations x up to 100 SQP iterations (+ slack)


EPS = 2.0 ** -52
FLOOR_SEEDS = 2


def _perturber(seed):
    """S -> S (1 +/- 2^-52) with a symmetric random sign pattern: the 1-ulp perturbation of scripts/parity_floor.py."""
    rng = np.random.default_rng(seed)

    def f(Sd, So):
        sd = rng.integers(0, 2, Sd.shape) * 2 - 1
        sd = np.triu(sd) + np.transpose(np.triu(sd, 1), (0, 2, 1))
        so = rng.integers(0, 2, So.shape) * 2 - 1
        return Sd * (1.0 + EPS * sd), So * (1.0 + EPS * so)
    return f


def _solve(args):
    xg, seed = args
    import bench
    from oracle import sqp
    model, c, cons = bench._oracle_problem(True)
    c.xg = np.asarray(xg)
    N = bench.N_KNOTS
    opts = dict(bench.SOLVER_OPTS)
    if seed >= 0:
        opts["_perturb_S"] = _perturber(seed)
    r = sqp.sqp(model, c, cons, np.zeros((12, N)), np.zeros((6, N - 1)), N, bench.DT, "PCG-SS", opts)
    outer_J = [row["J"] for row in r["trace"] if row["D"] is None]
    # QP solves per outer iteration: every QP solve appends exactly one row (accepted or failed search) after the outer-start row
    outer_qp, q = [], 0
    for row in r["trace"]:
        if row["D"] is None:
            outer_qp.append(0)
        else:
            outer_qp[-1] += 1
    seg = np.cumsum([0] + outer_qp)
    outer_pcg = [int(sum(r["pcg_iters"][seg[i]:seg[i + 1]])) for i in range(len(outer_qp))]
    outer_ls = [int(sum(r["ls_trials"][seg[i]:seg[i + 1]])) for i in range(len(outer_qp))]
    return dict(exits=[r["exit_sqp"], r["exit_soft"], r["outer_iter"], r["sqp_iter"]], pcg=r["pcg_iters"], ls=r["ls_trials"],
                J=r["J"], c=r["c"], x=r["x"], u=r["u"], outer_J=outer_J, outer_qp=outer_qp, outer_pcg=outer_pcg, outer_ls=outer_ls)


def main():
    import multiprocessing as mp
    import bench
    n_inst = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    xg = bench.workload_goals(1, 0, 8192)[:n_inst]
    t0 = time.perf_counter()
    with mp.get_context("fork").Pool(min(os.cpu_count() or 1, n_inst)) as pool:
        res_all = pool.map(_solve, [(g, s) for s in range(-1, FLOOR_SEEDS) for g in xg], chunksize=1)
    el = time.perf_counter() - t0
    res = res_all[:n_inst]
    # parity floor of the oracle itself on these instances: how far the result moves when S is perturbed by 1 ulp in every QP solve
    floor_same = np.zeros(n_inst, dtype=np.int32); floor_J = np.zeros(n_inst); floor_x = np.zeros(n_inst); floor_u = np.zeros(n_inst)
    for s in range(FLOOR_SEEDS):
        for i in range(n_inst):
            p, b = res_all[(s + 1) * n_inst + i], res[i]
            same = p["exits"] == b["exits"] and p["pcg"] == b["pcg"] and p["ls"] == b["ls"]
            floor_same[i] += int(same)
            if same:
                floor_J[i] = max(floor_J[i], abs(p["J"] - b["J"]) / max(1.0, abs(b["J"])))
                floor_x[i] = max(floor_x[i], float(np.max(np.abs(p["x"] - b["x"]))))
                floor_u[i] = max(floor_u[i], float(np.max(np.abs(p["u"] - b["u"]))))
    pcg_seq = -np.ones((n_inst, QP_CAP), dtype=np.int16)
    ls_seq = -np.ones((n_inst, QP_CAP), dtype=np.int8)
    outer_J = np.full((n_inst, 10), np.nan)
    outer_qp = np.zeros((n_inst, 10), dtype=np.int32); outer_pcg = np.zeros((n_inst, 10), dtype=np.int64); outer_ls = np.zeros((n_inst, 10), dtype=np.int64)
    for i, r in enumerate(res):
        assert len(r["pcg"]) <= QP_CAP
        pcg_seq[i, :len(r["pcg"])] = r["pcg"]
        ls_seq[i, :len(r["ls"])] = r["ls"]
        outer_J[i, :len(r["outer_J"])] = r["outer_J"]
        outer_qp[i, :len(r["outer_qp"])] = r["outer_qp"]; outer_pcg[i, :len(r["outer_pcg"])] = r["outer_pcg"]; outer_ls[i, :len(r["outer_ls"])] = r["outer_ls"]
    out = dict(xg=xg, exits=np.array([r["exits"] for r in res], dtype=np.int32),
               qp=np.array([len(r["pcg"]) for r in res], dtype=np.int32),
               total_pcg=np.array([sum(r["pcg"]) for r in res], dtype=np.int64),
               total_trials=np.array([sum(r["ls"]) for r in res], dtype=np.int64),
               J=np.array([r["J"] for r in res]), c=np.array([r["c"] for r in res]),
               x=np.stack([r["x"] for r in res]), u=np.stack([r["u"] for r in res]),
               pcg_seq=pcg_seq, ls_seq=ls_seq, outer_J=outer_J, outer_qp=outer_qp, outer_pcg=outer_pcg, outer_ls=outer_ls,
               floor_seeds=np.int32(FLOOR_SEEDS), floor_counts_identical=floor_same, floor_rel_J=floor_J, floor_abs_x=floor_x, floor_abs_u=floor_u)
    path = os.path.join(HERE, "c4_fixture.npz")
    np.savez_compressed(path, **out)
    print("wrote %s: %d instances in %.1f s; qp solves per instance mean %.1f, exit_soft hist %s" %
          (path, n_inst, el, out["qp"].mean(), np.bincount(out["exits"][:, 1], minlength=4).tolist()))
    print("floor (1-ulp perturbation of S, %d seeds): %d of %d instances keep every count under all seeds; on those max rel J %.1e, max |dx| %.1e, max |du| %.1e" %
          (FLOOR_SEEDS, int((floor_same == FLOOR_SEEDS).sum()), n_inst, floor_J[floor_same == FLOOR_SEEDS].max(), floor_x[floor_same == FLOOR_SEEDS].max(),
           floor_u[floor_same == FLOOR_SEEDS].max()))
    for i in np.nonzero(floor_same < FLOOR_SEEDS)[0]:
        print("  instance %d (qp %d): counts change under a 1-ulp perturbation in %d of %d seeds" % (i, out["qp"][i], FLOOR_SEEDS - floor_same[i], FLOOR_SEEDS))


if __name__ == "__main__":
    main()
