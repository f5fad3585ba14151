"""-m gpu: recording of solver internals against what the reference's authors recorded in data/{3,4} (fixtures in
tests/golden/ref_data.npz, extracted from the reference's own pickles by make_golden.py)."""
import os

import numpy as np
import pytest

import trajoptmpcreference_b200 as t
from trajoptmpcreference_b200 import record
from conftest import load_npz
from gpu_common import make_pair

pytestmark = pytest.mark.gpu


def _vals(lst):
    return [np.asarray(e["value"]) for e in lst]


@pytest.mark.parametrize("run", ["4", "3"])
def test_saved_lists_match_author_recordings(run, oracle_models):
    D = load_npz("ref_data.npz")
    N = 10
    (plant, pc, _), _ = make_pair("arm2", N, oracle_models, xg=D[run + "/xg"])
    solver = t.TrajoptMPCReference(plant, pc)
    x, u, e1, e2, outer, it = solver.SQP(np.zeros((4, N)), np.zeros((2, N - 1)), N, 0.1, t.SQPSolverMethods.PCG_SS,
                                         options={"expected_reduction_min_SQP_DDP": -100, "overloading": False}, record=True)
    assert [e1, e2, outer, it] == D[run + "/exits"].tolist()
    n_it = D[run + "/G"].shape[0]
    assert len(solver.saved_G) == n_it == len(solver.saved_dxul) == len(solver.saved_Pinv) == len(solver.saved_invG)
    # later iterations inherit the parity floor of the iterate they are evaluated at; the first is exact to rounding
    for i in range(n_it):
        tol = 1e-12 if i == 0 else 2e-6
        for name in ("G", "g", "C", "c", "invG", "Pinv"):
            got = _vals(getattr(solver, "saved_" + name))[i]
            ref = D[run + "/" + name][i]
            scale = max(1.0, np.max(np.abs(ref)))
            assert got.shape == ref.shape, name
            assert np.max(np.abs(got - ref)) < tol * scale * (1e3 if name in ("invG", "Pinv") else 1), (name, i)
        # dxul carries the PCG exit tolerance (1e-6 on |r^T Pinv r|)
        assert np.max(np.abs(_vals(solver.saved_dxul)[i] - D[run + "/dxul"][i])) < 5e-3 * max(1.0, np.max(np.abs(D[run + "/dxul"][i])))
    assert [e["iteration"] for e in solver.saved_G] == list(range(n_it))
    assert len(solver.saved_x) == D[run + "/x"].shape[0]
    for got, ref in zip(_vals(solver.saved_x), D[run + "/x"]):
        assert np.max(np.abs(got - ref)) < 1e-4
    for got, ref in zip(_vals(solver.saved_u), D[run + "/u"]):
        assert np.max(np.abs(got - ref)) < 1e-4
    for k in range(N - 1):
        assert np.allclose(solver.saved_Ak[k]["value"], D[run + "/Ak_first"][k], rtol=1e-12, atol=1e-13)
        assert np.allclose(solver.saved_Bk[k]["value"], D[run + "/Bk_first"][k], rtol=1e-12, atol=1e-13)
        assert np.allclose(solver.saved_xkp1[k]["value"], D[run + "/xkp1_first"][k], rtol=1e-12, atol=1e-13)
    # the recorded blocks are consistent with each other: S = -C invG C^T, gamma = c - C invG g, and a recorded solve does not change
    # the result of the solve
    G0, C0, iG0 = _vals(solver.saved_G)[0], _vals(solver.saved_C)[0], _vals(solver.saved_invG)[0]
    assert np.max(np.abs(iG0 @ G0 - np.eye(G0.shape[0]))) < 1e-9
    assert np.max(np.abs(_vals(solver.saved_S)[0] + C0 @ iG0 @ C0.T)) < 1e-9 * np.max(np.abs(_vals(solver.saved_S)[0]))
    assert np.max(np.abs(_vals(solver.saved_gamma)[0] - (_vals(solver.saved_c)[0] - C0 @ iG0 @ _vals(solver.saved_g)[0]))) < 1e-9
    x2, u2, *_ = solver.SQP(np.zeros((4, N)), np.zeros((2, N - 1)), N, 0.1, t.SQPSolverMethods.PCG_SS,
                            options={"expected_reduction_min_SQP_DDP": -100, "overloading": False})
    assert np.array_equal(x, x2) and np.array_equal(u, u2)


def test_record_files(tmp_path, oracle_models):
    """runSolversSQP(record=True) writes the reference's file set; the CSVs and pickles read back with pandas like the reference's."""
    import pandas as pd
    N = 10
    limits = {"torque": ([2.0], [-2.0], "AUGMENTED_LAGRANGIAN")}
    (plant, pc, pcons), _ = make_pair("pend", 20, oracle_models, limits=limits)
    res = record.runSQPExample(plant, pc, pcons, 20, 0.1, [t.SQPSolverMethods.PCG_SS], options={"expected_reduction_min_SQP_DDP": -100},
                               n_test=7, record=True, data_dir=str(tmp_path))
    x, u = res[0][0], res[0][1]
    base = tmp_path / "7"
    names = set(os.listdir(base))
    assert {"final_traj.csv", "final_input.csv", "results.plk", "trace.plk", "G.plk", "g.plk", "C.plk", "c.plk", "invG.plk", "Pinv.plk",
            "dxul.plk", "x.plk", "u.plk", "Ak.plk", "Bk.plk", "xkp1.plk", "inner_traces.plk", "J_tot_constraints.plk"} <= names
    fx = pd.read_csv(base / "final_traj.csv", index_col=0).to_numpy()
    fu = pd.read_csv(base / "final_input.csv", index_col=0).to_numpy()
    assert np.allclose(fx, x, rtol=0, atol=1e-15) and np.allclose(fu, u, rtol=0, atol=1e-15)
    tr = pd.read_pickle(base / "trace.plk")
    assert {"iteration", "outer_iteration", "alpha", "rho", "J", "c", "merit"} <= set(tr.columns)
    G = pd.read_pickle(base / "G.plk")
    assert {"value", "iteration", "outer_iteration", "line_search_iteration"} <= set(G.columns)
    assert G["value"][0].shape == (3 * 19 + 2, 3 * 19 + 2)
    # every outer (soft-constraint) iteration was recorded; the returned counter is one past the last one when the loop ends on
    # exit_soft == 1 (TrajoptMPCReference.py:483-508 increments before leaving)
    outers = sorted(set(int(v) for v in G["outer_iteration"]))
    assert outers == list(range(len(outers))) and outers[-1] in (res[0][4], res[0][4] - 1)
    results = pd.read_pickle(base / "results.plk")
    assert len(results) == 10
