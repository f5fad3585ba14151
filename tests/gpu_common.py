import json
import os

import numpy as np

import trajoptmpcreference_b200 as t
from conftest import GOLDEN
from oracle import cost as ocost, constraint as ocons


def make_pair(name, N, oracle_models, xg=None, limits=None, integrator=0, cost_kind=None):
    """(product objects, oracle objects) for the standard weights Q=I, QF=100I, R=0.1I."""
    m = oracle_models[name]
    n = m.n; nx = 2 * n
    Q, QF, R = np.eye(nx), 100.0 * np.eye(nx), 0.1 * np.eye(n)
    if cost_kind is None:
        cost_kind = "urdf" if name == "arm2" else "quadratic"
    if xg is None:
        xg = np.array([-1.0, 1.5, 0, 0]) if cost_kind == "urdf" else (np.concatenate([np.linspace(0.5, -0.5, n), np.zeros(n)]) if name != "pend" else np.array([3.14159, 0.0]))
    plant = t.URDFPlant(integrator_type=integrator, options={"path_to_urdf": name})
    if cost_kind == "urdf":
        Q, QF = np.eye(4), 100.0 * np.eye(4)       # weights of the end-effector state (x, y, vx, vy)
        pc = t.UrdfCost(plant, Q.copy(), QF.copy(), R.copy(), np.array(xg, dtype=float))
        oc = ocost.UrdfCost(m, Q, QF, R, xg)
    else:
        pc = t.QuadraticCost(Q.copy(), QF.copy(), R.copy(), np.array(xg, dtype=float))
        oc = ocost.QuadraticCost(Q, QF, R, xg)
    pcons, ocn = None, None
    if limits:
        pcons = t.TrajoptConstraint(n, n, n, N)
        ocn = ocons.SoftConstraints(n, n, n, N)
        for kind, (ub, lb, mode) in limits.items():
            getattr(pcons, "set_%s_limits" % kind)(list(ub), list(lb), mode, {})
            getattr(ocn, "set_%s_limits" % kind)(ub, lb, mode)
    return (plant, pc, pcons), (m, oc, ocn)


def solve_meta():
    with open(os.path.join(GOLDEN, "solve_meta.json")) as f:
        return json.load(f)
