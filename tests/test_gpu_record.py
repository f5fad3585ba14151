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


@pytest.mark.parametrize("run", ["4", "3"])
def test_cost_and_plant_level_lists_match_author_recordings(run, oracle_models):
    """UrdfCost.saved_cost / saved_dx / saved_grad / saved_hess / saved_Jacobian_tot_state (TrajoptCost.py:411-517) and
    URDFPlant.saved_Minv / saved_qdd / saved_dc_du / saved_dqdd (TrajoptPlant.py:297-322): one entry per callback invocation of the
    reference, reproduced by replaying its call sequence on the recorded iterates (record.py).  Entry counts, the iteration /
    line-search tags and the values against the authors' pickles of data/3, data/4 (first QP solve exact to rounding; later entries
    inherit the parity floor of the iterate they are evaluated at)."""
    D = load_npz("ref_data.npz")
    N = 10
    (plant, pc, _), _ = make_pair("arm2", N, oracle_models, xg=D[run + "/xg"])
    solver = t.TrajoptMPCReference(plant, pc)
    solver.SQP(np.zeros((4, N)), np.zeros((2, N - 1)), N, 0.1, t.SQPSolverMethods.PCG_SS,
               options={"expected_reduction_min_SQP_DDP": -100, "overloading": False}, record=True)
    post = 4 * (N - 1) + 2          # cost.value calls of runSolversSQP's post-processing (exampleHelpers.py:98-107), recorded after the solve
    lists = {"cost": pc.saved_cost, "dx": pc.saved_dx, "grad": pc.saved_grad, "hess": pc.saved_hess, "Jacobian_tot_state": pc.saved_Jacobian_tot_state,
             "Minv": plant.saved_Minv, "qdd": plant.saved_qdd, "dc_du": plant.saved_dc_du, "dqdd": plant.saved_dqdd}
    floor_x = 10 * {"4": 1.7e-7, "3": 4.1e-6}[run]          # tests/golden/floor.json (arm2_N10_SS, arm2_N10_SS_xg3)
    for name, got in lists.items():
        ref, tags = D["%s/lvl_%s" % (run, name)], D["%s/lvl_%s_tags" % (run, name)]
        n_ref = len(ref) - (post if name in ("cost", "dx") else 0)
        assert len(got) == n_ref, (name, len(got), n_ref)
        if name not in ("cost", "dx"):      # the authors' build did not tag cost.value calls (all zeros in cost.plk)
            assert [[e["iteration"], e["outer_iteration"], e["line_search_iteration"]] for e in got] == tags[:n_ref].tolist(), name
        first = {"cost": N, "dx": N, "grad": N, "hess": N, "Jacobian_tot_state": 2 * N, "Minv": 4 * (N - 1), "qdd": 4 * (N - 1), "dc_du": N - 1, "dqdd": N - 1}[name]
        worst = 0.0
        for i, e in enumerate(got):
            v = np.asarray(e["value"], dtype=float).reshape(-1)
            r = ref[i][:v.size]
            assert not np.any(np.isnan(r)) and (ref[i].size == v.size or np.all(np.isnan(ref[i][v.size:]))), (name, i)
            scale = max(1.0, float(np.max(np.abs(r))))
            err = float(np.max(np.abs(v - r))) / scale
            worst = max(worst, err)
            assert err < (1e-12 if i < first else 1e3 * floor_x), (name, i, err)
        print("data/%s %-20s %4d entries, worst relative error %.1e" % (run, name, len(got), worst))
    # direct calls after the solve keep appending while recording is on, like the reference's post-processing
    n0 = len(pc.saved_cost)
    pc.value(np.zeros(4), np.zeros(2)); pc.value(np.zeros(4), None)
    assert len(pc.saved_cost) == n0 + 2 and len(pc.saved_dx) == n0 + 2


@pytest.mark.parametrize("integ", [1, 2, 3])
def test_plant_level_lists_follow_the_integrator_branch(integ, oracle_models):
    """URDFPlant.saved_Minv / saved_qdd / saved_dc_du / saved_dqdd of a complete solve with the semi-implicit (1), midpoint (2) and rk3 (3)
    integrators against lists recorded from the unmodified reference (tests/golden/record_integrators.npz): the multi-stage branches
    call forward_dynamics / forward_dynamics_gradient once per stage point (TrajoptPlant.py:141-156, :171-198), the semi-implicit
    branch drops the counters on the gradient call (:131).  Counts, counters and values."""
    D = load_npz("record_integrators.npz")
    N = 6
    (plant, pc, _), _ = make_pair("pend", N, oracle_models, xg=D["%d/xg" % integ], integrator=integ)
    solver = t.TrajoptMPCReference(plant, pc)
    x, u, e1, e2, outer, it = solver.SQP(np.zeros((2, N)), np.zeros((1, N - 1)), N, 0.1, t.SQPSolverMethods.PCG_SS,
                                         options={"expected_reduction_min_SQP_DDP": -100, "overloading": False}, record=True)
    assert [int(e1), int(e2), int(outer), int(it)] == D["%d/exits" % integ].tolist()
    assert np.max(np.abs(x - D["%d/x" % integ])) < 1e-9 and np.max(np.abs(u - D["%d/u" % integ])) < 1e-9
    for name in ("Minv", "qdd", "dc_du", "dqdd"):
        got = getattr(plant, "saved_" + name)
        ref, tags = D["%d/lvl_%s" % (integ, name)], D["%d/lvl_%s_tags" % (integ, name)]
        assert len(got) == len(ref), (name, len(got), len(ref))
        assert [[e["iteration"], e["outer_iteration"], e["line_search_iteration"]] for e in got] == tags.tolist(), name
        worst = max(float(np.max(np.abs(np.asarray(e["value"], dtype=float).reshape(-1) - r))) / max(1.0, float(np.max(np.abs(r)))) for e, r in zip(got, ref))
        assert worst < 1e-9, (name, worst)
