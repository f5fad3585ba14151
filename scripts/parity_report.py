#!/usr/bin/env python
"""Achieved GPU-vs-reference errors on every golden solve case, next to the measured parity floor (GPU box).

  python scripts/parity_report.py            ->  gpurun_out/parity_report.json  (+ a table on stdout)

For each case of tests/golden/solve.npz (outputs of the unmodified reference, tests/golden/make_golden.py): exits and iteration
counts identical?, relative error of J, max |dx|, |du|, multipliers; beside them the floor of tests/golden/floor.json (how far
the reference's own result moves under a 1-ulp perturbation of S, scripts/parity_floor.py) and the north-star 1e-9 target.
For the cases with soft limits the first outer iteration / trace row at which the GPU path leaves the reference's path is
located by re-solving with max_iter_softConstraints = 1, 2, ...
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import trajoptmpcreference_b200 as t          # noqa: E402
from conftest import load_npz                   # noqa: E402
from gpu_common import make_pair, solve_meta    # noqa: E402
from oracle import rbd                          # noqa: E402


def main():
    with open(os.path.join(ROOT, "tests", "golden", "models.json")) as f:
        models = {k: rbd.Model(v) for k, v in json.load(f).items()}
    with open(os.path.join(ROOT, "tests", "golden", "floor.json")) as f:
        floor = json.load(f)["cases"]
    S = load_npz("solve.npz")
    rows = {}
    for tag, mt in solve_meta().items():
        N = mt["N"]
        (plant, pc, pcons), _ = make_pair(mt["robot"], N, models, xg=S[tag + "/xg"], limits=mt["limits"], integrator=mt["integrator"])
        solver = t.TrajoptMPCReference(plant, pc, pcons) if pcons is not None else t.TrajoptMPCReference(plant, pc)
        n = plant.get_num_pos()
        opts = dict(mt["options"]); opts["overloading"] = False
        x, u, e1, e2, outer, it = solver.SQP(np.zeros((2 * n, N)), np.zeros((n, N - 1)), N, 0.1, getattr(t.SQPSolverMethods, mt["method"]), options=opts)
        ref_exits = S[tag + "/exits"].tolist()
        k = len(solver.trace) - 1
        pcg_same = True
        if mt["method"].startswith("PCG"):
            pcg_same = [r["pcg_iters"] for r in solver.trace[1:]] == S[tag + "/pcg_iters"].tolist()[-k:] if k else True
        total_qp = int(solver.last_result.total_qp[0]); total_pcg = int(solver.last_result.total_pcg[0])
        row = {"exits_gpu": [int(e1), int(e2), int(outer), int(it)], "exits_ref": ref_exits, "exits_identical": [int(e1), int(e2), int(outer), int(it)] == ref_exits,
               "pcg_counts_last_outer_identical": bool(pcg_same), "qp_solves_gpu": total_qp, "pcg_total_gpu": total_pcg,
               "pcg_total_ref": int(S[tag + "/pcg_iters"].sum()), "qp_solves_ref": int(len(S[tag + "/pcg_iters"])) if mt["method"].startswith("PCG") else int(len(S[tag + "/tr_ls"])),
               "rel_J": abs(float(solver.last_result.J[0]) - float(S[tag + "/J"])) / max(1e-300, abs(float(S[tag + "/J"]))),
               "abs_x": float(np.max(np.abs(x - S[tag + "/x"]))), "abs_u": float(np.max(np.abs(u - S[tag + "/u"]))),
               "rel_x": float(np.max(np.abs(x - S[tag + "/x"])) / max(1e-300, np.max(np.abs(S[tag + "/x"])))),
               "rel_u": float(np.max(np.abs(u - S[tag + "/u"])) / max(1e-300, np.max(np.abs(S[tag + "/u"])))),
               "floor": floor.get(tag)}
        if pcons is not None and pcons.torque_limits.is_soft_constraint_mode() and mt["method"].startswith("PCG"):
            row["abs_mu"] = float(np.max(np.abs(pcons.torque_limits.quadratic_penalty_mu - S[tag + "/mu"]) / np.maximum(1e-300, np.abs(S[tag + "/mu"]))))
            row["abs_lam"] = float(np.max(np.abs(pcons.torque_limits.augmented_lagrangian_lambda - S[tag + "/lam"])))
            # locate the first outer iteration whose cumulative QP / PCG counts differ from the reference's trace
            tr_outer = S[tag + "/tr_outer"]; pcg_ref = S[tag + "/pcg_iters"]
            first_bad = None
            for o in range(1, int(ref_exits[2]) + 2):
                (plant2, pc2, pcons2), _ = make_pair(mt["robot"], N, models, xg=S[tag + "/xg"], limits=mt["limits"], integrator=mt["integrator"])
                s2 = t.TrajoptMPCReference(plant2, pc2, pcons2)
                o2 = dict(opts); o2["max_iter_softConstraints"] = o
                s2.SQP(np.zeros((2 * n, N)), np.zeros((n, N - 1)), N, 0.1, getattr(t.SQPSolverMethods, mt["method"]), options=o2)
                sel = tr_outer < o
                want = (int(sel.sum()), int(pcg_ref[:len(tr_outer)][sel].sum()))
                got = (int(s2.last_result.total_qp[0]), int(s2.last_result.total_pcg[0]))
                if got != want:
                    first_bad = {"outer_iterations": o, "gpu_qp_pcg": got, "ref_qp_pcg": want}
                    break
            row["first_divergent_outer_iteration"] = first_bad
        rows[tag] = row
        fl = floor.get(tag, {})
        print("%-22s exits %s pcg %s  rel J %.1e (floor %.1e)  |dx| %.1e (floor %.1e)  |du| %.1e (floor %.1e)%s" %
              (tag, "same" if row["exits_identical"] else "DIFF %s vs %s" % (row["exits_gpu"], ref_exits), "same" if pcg_same else "DIFF",
               row["rel_J"], fl.get("rel_J", float("nan")), row["abs_x"], fl.get("abs_x", float("nan")), row["abs_u"], fl.get("abs_u", float("nan")),
               "" if "first_divergent_outer_iteration" not in row else "  first divergent outer iteration: %s" % (row["first_divergent_outer_iteration"],)))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "parity_report.json"), "w") as f:
        json.dump(rows, f, indent=1)


if __name__ == "__main__":
    main()
