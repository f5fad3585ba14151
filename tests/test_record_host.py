"""Host side of the recording path (trajoptmpcreference_b200/record.py) without a GPU: the Recorder is driven by a stand-in solver
whose `fetch` serves the ORACLE's KKT blocks in the C-ABI's knot-major layouts; the dense matrices it assembles must equal what the
reference itself returned for the same problem (tests/golden/kkt.npz: G, g, C, c, invG, S, gamma, Pinv, dxul of
formKKTSystemBlocks / solveKKTSystem_Schur)."""
import enum

import numpy as np
import pytest

from conftest import load_npz
from oracle import cost as ocost, kkt
from trajoptmpcreference_b200 import _lib
from trajoptmpcreference_b200.record import Recorder, save_in_file


class _Method(enum.Enum):
    PCG_SS = "PCG-SS"
    PCG_BJ = "PCG-BJ"


class _FakeSolver:
    """Serves oracle blocks like BatchSolver.fetch / get_status / get_scalars / get_trajectory (batch 1)."""

    def __init__(self, model, cost, X, U, xs, rho, kind):
        N, nx = X.shape
        nu = U.shape[1]
        self.N, self.nx, self.nu, self.m = N, nx, nu, nx + nu
        self.has_limits = False
        blocks = kkt.form_blocks(model, cost, None, X, U, xs, 0.1)
        sch = kkt.schur(blocks, rho, nx)
        Pd, Po = kkt.preconditioner(sch["Sd"], sch["So"], kind)
        l, _ = kkt.pcg(sch["Sd"], sch["So"], sch["gamma"], Pd, Po)[:2]
        dz = kkt.recover(blocks, sch, l, nx)
        m = self.m
        AB = np.zeros((N, nx, m)); AB[:N - 1] = sch["AB"]
        So = np.zeros((N, nx, nx)); So[1:] = sch["So"]
        xkp1 = np.zeros((N, nx)); xkp1[:N - 1] = blocks["xkp1"]
        Upad = np.zeros((N, nu)); Upad[:N - 1] = U
        self.arr = {"kkt_hess": blocks["G"].reshape(N, -1), "g": blocks["g"], "AB": AB.reshape(N, -1), "x": X, "u": Upad, "xkp1": xkp1,
                    "Ghat": sch["Ghat"].reshape(N, -1), "Sd": sch["Sd"].reshape(N, -1), "So": So.reshape(N, -1), "Pd": Pd.reshape(N, -1),
                    "gamma": sch["gamma"], "l": l.reshape(N, nx), "dz": dz.reshape(N, m), "cost_grad": cost.gradients(X, U)}
        self.rho = rho
        self.X, self.U = X, U

    def fetch(self, name):
        if name == "nu_trace":
            out = np.zeros((1, 128)); out[0, 0] = 1.0; out[0, 1] = 1e-9
            return out
        return self.arr[name][None]

    def get_status(self):
        return np.array([[0, 0, 0, 0, 0, 0, 0, 1]], dtype=np.int32)

    def get_scalars(self):
        return np.array([[0.0, 0.0, 0.0, self.rho]])

    def get_trajectory(self):
        return self.X.T[None].copy(), self.U.T[None].copy()


class _Owner:
    pass


@pytest.mark.parametrize("tag,name,kind", [("arm3_qc", "arm3", "SS"), ("arm6_qc", "arm6", "BJ")])
def test_recorder_dense_layout_matches_reference(tag, name, kind, oracle_models):
    Kz = load_npz("kkt.npz")
    model = oracle_models[name]
    n = model.n; nx = 2 * n
    X = Kz[tag + "/x"].T.copy(); U = Kz[tag + "/u"].T.copy()
    cost = ocost.QuadraticCost(np.eye(nx), 100.0 * np.eye(nx), 0.1 * np.eye(n), Kz[tag + "/xg"])
    rho = 1e-3
    s = _FakeSolver(model, cost, X, U, Kz[tag + "/xs"], rho, kind)
    owner = _Owner()
    rec = Recorder(owner, s, _Method["PCG_" + kind], Kz[tag + "/xs"], {})
    rec(_lib.HOOK_LINSYS, 0)
    rec(_lib.HOOK_STEP, 0)
    nz = G_n = Kz[tag + "/G"].shape[0]
    got = {k: np.asarray(getattr(owner, "saved_" + k)[0]["value"]) for k in ("G", "g", "C", "c", "invG", "S", "gamma", "Pinv", "l", "dxul")}
    assert np.allclose(got["G"], Kz[tag + "/G"] + rho * np.eye(nz), rtol=0, atol=1e-12)      # solveKKTSystem_Schur stores G + rho I (:369)
    for k in ("g", "C", "c"):
        assert got[k].shape == Kz[tag + "/" + k].shape and np.allclose(got[k], Kz[tag + "/" + k], rtol=1e-12, atol=1e-12), k
    scale = np.max(np.abs(Kz[tag + "/invG"]))
    assert np.max(np.abs(got["invG"] - Kz[tag + "/invG"])) < 1e-11 * scale
    assert np.max(np.abs(got["S"] - Kz[tag + "/S"])) < 1e-11 * np.max(np.abs(Kz[tag + "/S"]))
    assert np.max(np.abs(got["gamma"] - Kz[tag + "/gamma"])) < 1e-10 * max(1.0, np.max(np.abs(Kz[tag + "/gamma"])))
    P = Kz[tag + "/Pinv_" + kind]
    assert np.max(np.abs(got["Pinv"] - P)) < 1e-9 * np.max(np.abs(P))
    ref_dxul = Kz["%s/dxul_PCG-%s" % (tag, kind)]
    assert got["dxul"].shape == ref_dxul.shape
    # dxul carries PCG's absolute exit test |r^T Pinv r| < 1e-6 (two PCG runs agree only to that level on arm6)
    assert np.max(np.abs(got["dxul"] - ref_dxul)) < 1e-4 * max(1.0, np.max(np.abs(ref_dxul)))
    assert len(owner.saved_Ak) == X.shape[0] - 1 and owner.saved_Ak[0]["value"].shape == (nx, nx) and owner.saved_Bk[0]["value"].shape == (nx, n)
    assert owner.saved_G[0]["iteration"] == 0 and owner.saved_G[0]["outer_iteration"] == 0
    assert len(owner.saved_inner_traces) == 1 and owner.saved_inner_traces[0][0][0] == [1.0, 1e-9]


def test_save_in_file_conventions(tmp_path):
    """exampleHelpers.save_in_file (:41-55): 'saved_' is dropped from the file name, directories are created, CSV keeps pandas' index."""
    import pandas as pd
    save_in_file(str(tmp_path / "3" / "saved_dxul.plk"), [{"value": np.ones((2, 1)), "iteration": 0}])
    save_in_file(str(tmp_path / "3" / "final_traj.csv"), np.arange(6.0).reshape(2, 3), csv=True)
    assert (tmp_path / "3" / "dxul.plk").exists()
    df = pd.read_pickle(tmp_path / "3" / "dxul.plk")
    assert list(df.columns) == ["value", "iteration"]
    back = pd.read_csv(tmp_path / "3" / "final_traj.csv", index_col=0).to_numpy()
    assert np.array_equal(back, np.arange(6.0).reshape(2, 3))
