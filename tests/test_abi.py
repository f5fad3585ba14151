"""The C-ABI libraries build (nvcc cross-compiles without a GPU), load, and export every symbol include/b2t.h declares.
No compute calls here; on a box without a GPU solver creation must fail loudly (no CPU fallback)."""
import ctypes

import numpy as np
import pytest

import trajoptmpcreference_b200 as t
from trajoptmpcreference_b200 import _lib


@pytest.mark.parametrize("name", ["arm2", "arm6"])
def test_library_exports_every_declared_symbol(name):
    lib = _lib.load_builtin(name)
    decl = _lib.declared_symbols()
    assert len(decl) >= 30
    for sym in decl:
        assert hasattr(lib, sym), sym
    assert set(decl) == set(_lib._SIGNATURES), set(decl) ^ set(_lib._SIGNATURES)
    assert lib.b2t_abi_version() == 2
    assert lib.b2t_model_name().decode() == name
    nq, nx, nu = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    lib.b2t_model_dims(ctypes.byref(nq), ctypes.byref(nx), ctypes.byref(nu))
    n = {"arm2": 2, "arm6": 6}[name]
    assert (nq.value, nx.value, nu.value) == (n, 2 * n, n)
    o = _lib.Options()
    lib.b2t_default_options(ctypes.byref(o))
    assert o.max_iter_linSys == 100 and o.rho_init == 1e-3 and o.alpha_min == 0.005 and o.merit_mu == 10


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    plant = t.URDFPlant(options={"path_to_urdf": "arm2"})
    cost = t.QuadraticCost(np.eye(4), np.eye(4), np.eye(2), np.zeros(4))
    with pytest.raises(t.B2TError):
        t.BatchSolver(plant, cost, None, 8, 0.1)


def test_api_argument_errors():
    plant = t.URDFPlant(options={"path_to_urdf": "arm2"})
    with pytest.raises(ValueError):
        t.URDFPlant(integrator_type=7, options={"path_to_urdf": "arm2"})
    with pytest.raises(ValueError):
        t.URDFPlant(options={})
    with pytest.raises(ValueError):
        t.TrajoptMPCReference(plant, object())
    c = t.TrajoptConstraint(2, 2, 2, 10)
    with pytest.raises(ValueError):
        c.set_torque_limits([1.0], [-1.0], "FULL_SET")          # singular KKT in the reference itself
    with pytest.raises(ValueError):
        c.set_torque_limits([1.0], [-1.0], "ADMM_PROJECTION")   # "[!] ERROR NOT IMPLEMENTED YET" in the reference
    c.set_torque_limits([1.0], [-1.0], "ACTIVE_SET")            # hard mode, exact methods N / S
    assert c.torque_limits.is_hard_constraint_mode() and not c.torque_limits.is_soft_constraint_mode()
    with pytest.raises(ValueError):
        c.set_torque_limits([1.0, 2.0, 3.0], [-1.0], "QUADRATIC_PENALTY")
    c.set_torque_limits([1.0], [-1.0], "QUADRATIC_PENALTY")
    assert c.torque_limits.quadratic_penalty_mu.shape == (4, 9)
    mu, lam, phi = c.pack(10)
    assert mu.shape == (12, 10) and mu[4, 0] == 1e-2 and mu[10, 3] == 1e-2
    # options dict is filled in place like the reference (TrajoptMPCReference.py:91-115)
    s = t.TrajoptMPCReference(plant, t.QuadraticCost(np.eye(4), np.eye(4), np.eye(2), np.zeros(4)))
    o = {}
    s.set_default_options(o)
    assert o["rho_init_SQP_DDP"] == 0.001 and o["max_iter_softConstraints"] == 10
