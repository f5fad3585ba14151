#!/usr/bin/env python
"""Parity floor of the reference itself.

  python scripts/parity_floor.py [seeds=4]        ->  tests/golden/floor.json        (build container only: needs /root/reference)

PCG stops at an ABSOLUTE |r^T Pinv r| < 1e-6 and the SQP iteration amplifies rounding, so "<= 1e-9 relative" can lie below what the
reference's own arithmetic determines.  This script measures that: every golden solve case of tests/golden/make_golden.py is
re-run with the UNMODIFIED reference (tests/ref/refshim.py), except that the Schur complement handed to the linear solver is
perturbed by ONE ULP (each entry multiplied by 1 +/- 2^-52, symmetric random sign pattern) -- the size of error any other
summation order, BLAS build or FMA contraction introduces.  Recorded per case: whether the iteration counts survive, and how far
the final x, u, J move (max over the seeds).  Tests then assert `error <= max(10 x floor, 1e-9 relative)` instead of a blanket
tolerance, and DESIGN.md tabulates floor vs achieved.

Injection points (no reference file is modified): module global `TrajoptMPCReference.PCG` (TrajoptMPCReference.py:8, used at
:438) for the PCG methods, and `np.linalg.solve` (used at :354, :432) for the exact methods N / S.
"""
import io
import json
import os
import sys
import contextlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "ref"))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
sys.path.insert(0, ROOT)
os.environ.setdefault("OMP_NUM_THREADS", "1")
os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")

EPS = 2.0 ** -52


def _perturb(S, rng):
    s = rng.integers(0, 2, S.shape) * 2 - 1
    s = np.triu(s) + np.triu(s, 1).T          # symmetric sign pattern: S stays symmetric
    return S * (1.0 + EPS * s)


def _run(args):
    tag, seed = args
    import refshim
    import make_golden as mg
    R = refshim.load()
    import TrajoptMPCReference as TM
    case = [c for c in mg.SOLVE_CASES if c[0] == tag][0]
    _, name, N, meth, opts, limits, xg = case
    integ = mg.integ_of(tag)
    plant, cost, cons, solver, xg = mg.make_problem(R, name, N, limits=limits, integrator=integ, xg=xg)
    n = plant.get_num_pos()
    rng = np.random.default_rng(1000 + seed)
    orig_pcg, orig_solve = TM.PCG, np.linalg.solve
    if seed >= 0:
        class PerturbedPCG(orig_pcg):
            def __init__(self, A, b, *a, **k):
                super().__init__(_perturb(np.asarray(A, dtype=float), rng), b, *a, **k)
        TM.PCG = PerturbedPCG

        def solve(A, b):
            A = np.asarray(A, dtype=float)
            return orig_solve(_perturb(A, rng) if A.shape[0] == A.shape[1] and A.shape[0] > 3 * n else A, b)
        np.linalg.solve = solve
    o = dict(opts); o["overloading"] = False
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            x, u, e1, e2, outer, it = solver.SQP(np.zeros((2 * n, N)), np.zeros((n, N - 1)), N, 0.1, getattr(R.SQPSolverMethods, meth), options=o)
            J = float(solver.totalCost(x, u, N))
    finally:
        TM.PCG, np.linalg.solve = orig_pcg, orig_solve
    pcg = [len(t[0][0]) - 1 for t in solver.saved_inner_traces]
    return tag, seed, dict(x=np.array(x), u=np.array(u), J=J, exits=[int(e1), int(e2), int(outer), int(it)], pcg=pcg, qp=len(solver.saved_l))


def main():
    import multiprocessing as mp
    import make_golden as mg
    seeds = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    fpath = os.path.join(ROOT, "tests", "golden", "floor.json")
    have = {}
    if os.path.exists(fpath) and not os.environ.get("FORCE"):      # only the cases floor.json does not hold yet (FORCE=1: all)
        with open(fpath) as f:
            have = json.load(f)["cases"]
    tags = [c[0] for c in mg.SOLVE_CASES if c[0] not in have]
    jobs = [(t, s) for t in tags for s in range(-1, seeds)]          # seed -1 = unperturbed
    with mp.get_context("fork").Pool(os.cpu_count() or 1) as pool:
        res = pool.map(_run, jobs, chunksize=1)
    by = {}
    for tag, seed, r in res:
        by.setdefault(tag, {})[seed] = r
    out = dict(have)
    for tag in tags:
        base = by[tag][-1]
        fl = dict(rel_J=0.0, abs_x=0.0, abs_u=0.0, rel_x=0.0, rel_u=0.0, counts_identical=0, seeds=seeds, qp_solves=base["qp"], exits=base["exits"])
        for s in range(seeds):
            r = by[tag][s]
            same = r["exits"] == base["exits"] and r["pcg"] == base["pcg"]
            fl["counts_identical"] += int(same)
            if not same:
                continue
            fl["rel_J"] = max(fl["rel_J"], abs(r["J"] - base["J"]) / max(1e-300, abs(base["J"])))
            dx = float(np.max(np.abs(r["x"] - base["x"]))); du = float(np.max(np.abs(r["u"] - base["u"])))
            fl["abs_x"] = max(fl["abs_x"], dx); fl["abs_u"] = max(fl["abs_u"], du)
            fl["rel_x"] = max(fl["rel_x"], dx / max(1e-300, float(np.max(np.abs(base["x"])))))
            fl["rel_u"] = max(fl["rel_u"], du / max(1e-300, float(np.max(np.abs(base["u"])))))
        out[tag] = fl
        print("%-22s qp %4d  counts identical %d/%d  floor: rel J %.1e  |dx| %.1e (rel %.1e)  |du| %.1e (rel %.1e)" %
              (tag, fl["qp_solves"], fl["counts_identical"], seeds, fl["rel_J"], fl["abs_x"], fl["rel_x"], fl["abs_u"], fl["rel_u"]))
    with open(fpath, "w") as f:
        json.dump({"perturbation": "S * (1 +/- 2^-52), symmetric random signs, every QP solve", "cases": out}, f, indent=1)


if __name__ == "__main__":
    main()
