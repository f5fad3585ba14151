"""Generate the golden fixtures in this directory from the UNMODIFIED reference at /root/reference.

Run in the build container only (python tests/golden/make_golden.py [section ...]); the fixtures are committed,
the GPU box never needs the reference.  Sections: models dynamics integrators cost kkt solve refdata record_integrators
(`solve` only runs the cases that solve.npz does not hold yet; FORCE=1 regenerates all.)
"""
import json
import os
import sys
import io
import contextlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests", "ref"))
sys.path.insert(0, ROOT)
import refshim  # noqa: E402

URDF = os.path.join(ROOT, "trajoptmpcreference_b200", "urdf")
ROBOTS = ["pend", "arm1", "arm2", "arm3", "arm4", "arm6"]


def quiet():
    return contextlib.redirect_stdout(io.StringIO())


def ref_plant(R, name, integrator=0, gravity=-9.81):
    return R.URDFPlant(integrator_type=integrator,
                       options={"path_to_urdf": os.path.join(URDF, name + ".urdf"), "overloading": False, "gravity": gravity})


def model_from_reference(robot):
    """Dump the constants of the reference's Robot object (sympy matrices -> coefficient matrices)."""
    import sympy as sp
    n = robot.get_num_joints()
    out = {"name": robot.name, "n": n, "parent": [int(robot.get_parent_id(j)) for j in range(n)],
           "jtype": [], "S": [], "X0": [], "Xa": [], "Xb": [], "H0": [], "Ha": [], "Hb": [], "I": [], "damping": []}
    for j in range(n):
        joint = robot.get_joint_by_id(j)
        th = joint.theta
        out["jtype"].append(joint.jtype)
        out["S"].append([float(s) for s in joint.S])
        out["damping"].append(float(joint.damping))
        for key, M in (("X", joint.Xmat_sp), ("H", joint.Xmat_sp_hom)):
            M = sp.Matrix(M)
            r, c = M.shape
            m0 = np.zeros((r, c)); ma = np.zeros((r, c)); mb = np.zeros((r, c))
            for i in range(r):
                for k in range(c):
                    e = sp.expand(M[i, k])
                    if joint.jtype == "revolute":
                        ma[i, k] = float(e.coeff(sp.cos(th)))
                        mb[i, k] = float(e.coeff(sp.sin(th)))
                        m0[i, k] = float(e.subs({sp.cos(th): 0, sp.sin(th): 0}))
                    else:
                        ma[i, k] = float(e.coeff(th))
                        m0[i, k] = float(e.subs({th: 0}))
            # verify the affine reconstruction against the reference's own lambdified function
            f = sp.utilities.lambdify(th, M, "numpy")
            for t in (0.3, -1.7, 2.9):
                f1, f2 = (np.cos(t), np.sin(t)) if joint.jtype == "revolute" else (t, 0.0)
                assert np.max(np.abs(np.array(f(t), dtype=float) - (m0 + f1 * ma + f2 * mb))) < 1e-15
            out[key + "0"].append(m0.tolist()); out[key + "a"].append(ma.tolist()); out[key + "b"].append(mb.tolist())
        out["I"].append(np.array(robot.get_Imat_by_id(j), dtype=float).tolist())
    return out


def sec_models(R):
    models = {}
    for name in ROBOTS:
        plant = ref_plant(R, name)
        models[name] = model_from_reference(plant.robot)
    with open(os.path.join(HERE, "models.json"), "w") as f:
        json.dump(models, f)
    # the generated arm URDFs must be parsed by the reference exactly like its own models/armK.urdf
    for k in (2, 3, 4, 5):
        a = model_from_reference(R.URDFPlant(options={"path_to_urdf": os.path.join(refshim.REF, "models", "arm%d.urdf" % k), "overloading": False}).robot)
        b = model_from_reference(ref_plant(R, "arm%d" % k).robot)
        a.pop("name"); b.pop("name")
        assert json.dumps(a) == json.dumps(b), k
    print("models.json written:", list(models))


def kat_point(n):
    """GRiD/util/util.py:78-103 fixed 'C++ RNG' test point."""
    q = [-0.336899, 1.29662, -0.677475, -1.42182, -0.706676, -0.134981]
    qd = [0.43302, -0.421561, -0.645439, -1.86055, -0.0130938, -0.458284]
    u = [0.741788, 1.92844, -0.903882, 0.0333959, 1.17986, -1.94599]
    return np.array(q[:n]), np.array(qd[:n]), np.array(u[:n])


def integ_of(tag):
    """integrator type of a solve case, from its tag suffix"""
    for suffix, it in (("_semi", 1), ("_mid", 2), ("_rk3", 3)):
        if tag.endswith(suffix):
            return it
    return 0


def sec_integrators(R):
    """Integrator types 2 (midpoint), 3 (rk3), 4 (rk4) of the unmodified reference (TrajoptPlant.py:140-270) at the points of
    sec_dynamics: step and (A, B) for 2 / 3; for 4 the step only -- its gradient branch raises TypeError (:259), recorded as such."""
    out = {}
    rng = np.random.default_rng(1234)
    for name in ROBOTS:
        plants = {it: ref_plant(R, name, it) for it in (2, 3, 4)}
        n = plants[2].get_num_pos()
        pts = [kat_point(n)] + [(rng.uniform(-2, 2, n), rng.uniform(-2, 2, n), rng.uniform(-2, 2, n)) for _ in range(5)]
        rec = {k: [] for k in ("q", "qd", "u", "A2", "B2", "xn2", "A3", "B3", "xn3", "xn4")}
        for q, qd, u in pts:
            x = np.concatenate([q, qd])
            for it in (2, 3):
                A, B = plants[it].integrator(x, u, 0.1, return_gradient=True)
                rec["A%d" % it].append(np.array(A)); rec["B%d" % it].append(np.array(B))
                rec["xn%d" % it].append(np.array(plants[it].integrator(x, u, 0.1)).reshape(-1))
            rec["xn4"].append(np.array(plants[4].integrator(x, u, 0.1)).reshape(-1))
            try:
                plants[4].integrator(x, u, 0.1, return_gradient=True)
                raised = ""
            except Exception as e:      # noqa: BLE001
                raised = type(e).__name__
            for k, val in (("q", q), ("qd", qd), ("u", u)):
                rec[k].append(np.array(val, dtype=float))
        for k, lst in rec.items():
            out[name + "/" + k] = np.stack(lst)
        out[name + "/rk4_gradient_raises"] = np.array(raised)
        print(name, "rk4 gradient raises:", raised)
    np.savez_compressed(os.path.join(HERE, "integrators.npz"), **out)
    print("integrators.npz written")


def sec_dynamics(R):
    """Per robot: the fixed known-answer point + seeded random points -> every dynamics quantity on the path."""
    out = {}
    rng = np.random.default_rng(1234)
    for name in ROBOTS:
        plants = {it: ref_plant(R, name, it) for it in (0, 1)}
        plant = plants[0]
        n = plant.get_num_pos()
        pts = [kat_point(n)] + [(rng.uniform(-2, 2, n), rng.uniform(-2, 2, n), rng.uniform(-2, 2, n)) for _ in range(5)]
        rec = {k: [] for k in ("q", "qd", "u", "c", "Minv", "qdd", "dc_du", "dqdd", "A0", "B0", "xn0", "A1", "B1", "xn1", "v", "a", "f")}
        if n == 2:
            rec.update({k: [] for k in ("ee", "J", "Jtot")})
        for q, qd, u in pts:
            x = np.concatenate([q, qd])
            rb = plant.rbdReference
            c = rb.rnea(q, qd, None, -9.81)[0]
            Minv = rb.minv(q)
            qdd = Minv @ (u - c)
            _, v, a, f = rb.rnea(q, qd, qdd, -9.81)
            dc_du = rb.rnea_grad(q, qd, qdd, -9.81)
            dqdd = plant.forward_dynamics_gradient(x, u)
            for it in (0, 1):
                A, B = plants[it].integrator(x, u, 0.1, return_gradient=True)
                xn = plants[it].integrator(x, u, 0.1)
                rec["A%d" % it].append(np.array(A)); rec["B%d" % it].append(np.array(B)); rec["xn%d" % it].append(np.array(xn).reshape(-1))
            for k, val in (("q", q), ("qd", qd), ("u", u), ("c", c), ("Minv", Minv), ("qdd", qdd), ("dc_du", dc_du),
                           ("dqdd", dqdd), ("v", v.T), ("a", a.T), ("f", f.T)):
                rec[k].append(np.array(val, dtype=float))
            if n == 2:
                rec["ee"].append(np.array(rb.end_effector_positions(q)).reshape(-1))
                rec["J"].append(np.array(rb.Jacobian(q)))
                rec["Jtot"].append(np.array(rb.jacobian_tot_state(q, qd)))
        for k, lst in rec.items():
            out[name + "/" + k] = np.stack(lst)
    np.savez_compressed(os.path.join(HERE, "dynamics.npz"), **out)
    print("dynamics.npz written")


def make_problem(R, name, N, seed=None, limits=None, cost_kind=None, integrator=0, xg=None):
    """Reference-side problem objects.  limits = dict(kind -> (upper, lower, mode))."""
    plant = ref_plant(R, name, integrator)
    n = plant.get_num_pos()
    nx = 2 * n
    Q = np.eye(nx); QF = 100.0 * np.eye(nx); Rm = 0.1 * np.eye(n)
    if cost_kind is None:
        cost_kind = "urdf" if name == "arm2" else "quadratic"
    if cost_kind == "urdf":
        xg = np.array([-1.0, 1.5, 0.0, 0.0]) if xg is None else np.asarray(xg, float)
        cost = R.UrdfCost(plant, Q, QF, Rm, xg)
    else:
        if xg is None:
            xg = np.concatenate([np.linspace(0.5, -0.5, n), np.zeros(n)]) if name != "pend" else np.array([3.14159, 0.0])
        cost = R.QC(Q, QF, Rm, np.asarray(xg, float))
    cons = None
    if limits:
        cons = R.TrajoptConstraint(n, n, n, N)
        for kind, (ub, lb, mode) in limits.items():
            getattr(cons, "set_%s_limits" % kind)(list(ub), list(lb), mode, options={"overloading": False})
    solver = R.TrajoptMPCReference(plant, cost, cons) if cons is not None else R.TrajoptMPCReference(plant, cost)
    return plant, cost, cons, solver, xg


KKT_CASES = [
    # tag, robot, N, limits
    ("arm2_urdf", "arm2", 6, None),
    ("arm3_qc", "arm3", 5, None),
    ("arm6_qc", "arm6", 4, None),
    ("pend_al", "pend", 6, {"torque": ([0.3], [-0.3], "AUGMENTED_LAGRANGIAN")}),
]


def sec_kkt(R):
    """One QP solve at a seeded random trajectory, every intermediate of solveKKTSystem_Schur, all 5 methods."""
    out = {}
    rng = np.random.default_rng(7)
    for tag, name, N, limits in KKT_CASES:
        plant, cost, cons, solver, xg = make_problem(R, name, N, limits=limits)
        n = plant.get_num_pos(); nx = 2 * n
        x = rng.uniform(-0.8, 0.8, (nx, N)); u = rng.uniform(-0.8, 0.8, (n, N - 1))
        xs = x[:, 0] + rng.uniform(-0.1, 0.1, nx)
        if cons is not None:   # non-trivial multipliers
            cons.torque_limits.augmented_lagrangian_lambda[:] = rng.uniform(-0.05, 0.05, cons.torque_limits.augmented_lagrangian_lambda.shape)
            cons.torque_limits.quadratic_penalty_mu[:] = rng.uniform(0.5, 2.0, cons.torque_limits.quadratic_penalty_mu.shape)
            out[tag + "/lam"] = cons.torque_limits.augmented_lagrangian_lambda.copy()
            out[tag + "/mu"] = cons.torque_limits.quadratic_penalty_mu.copy()
        rho = 1e-3
        out[tag + "/x"] = x; out[tag + "/u"] = u; out[tag + "/xs"] = xs; out[tag + "/xg"] = xg
        with quiet():
            G, g, C, c = solver.formKKTSystemBlocks(x, u, xs, N, 0.1)
            J = solver.totalCost(x, u, N)
            cv = solver.totalHardConstraintViolation(x, u, xs, N, 0.1)
        out[tag + "/G"] = G; out[tag + "/g"] = g; out[tag + "/C"] = C; out[tag + "/c"] = c
        out[tag + "/J"] = np.float64(J); out[tag + "/cv"] = np.float64(cv)
        with quiet():
            dxul = solver.solveKKTSystem(x, u, xs, N, 0.1, rho, {})
        out[tag + "/dxul_N"] = dxul
        for meth in ("S", "PCG-J", "PCG-BJ", "PCG-SS"):
            opts = {"DEBUG_MODE": False}
            use_pcg = meth != "S"
            if use_pcg:
                opts.update({"exit_tolerance": 1e-6, "max_iter": 100, "RETURN_TRACE": False, "preconditioner_type": meth[4:]})
            solver.saved_S.clear(); solver.saved_gamma.clear(); solver.saved_Pinv.clear(); solver.saved_inner_traces.clear(); solver.saved_l.clear(); solver.saved_invG.clear()
            with quiet():
                dxul = solver.solveKKTSystem_Schur(x, u, xs, N, 0.1, rho, use_pcg, opts)
            out["%s/dxul_%s" % (tag, meth)] = dxul
            out["%s/l_%s" % (tag, meth)] = np.array(solver.saved_l[-1]["value"])
            if meth == "S":
                out[tag + "/S"] = solver.saved_S[-1]["value"]; out[tag + "/gamma"] = solver.saved_gamma[-1]["value"]
                out[tag + "/invG"] = solver.saved_invG[-1]["value"]
            else:
                out["%s/Pinv_%s" % (tag, meth[4:])] = solver.saved_Pinv[-1]["value"]
                out["%s/trace_%s" % (tag, meth[4:])] = np.array(solver.saved_inner_traces[-1][0][0])
    np.savez_compressed(os.path.join(HERE, "kkt.npz"), **out)
    print("kkt.npz written")


SOLVE_CASES = [
    # tag, robot, N, method, options, limits, xg
    ("arm2_N10_SS", "arm2", 10, "PCG_SS", {"expected_reduction_min_SQP_DDP": -100}, None, None),
    ("arm2_N10_BJ", "arm2", 10, "PCG_BJ", {"expected_reduction_min_SQP_DDP": -100}, None, None),
    ("arm2_N10_S", "arm2", 10, "S", {"expected_reduction_min_SQP_DDP": -100}, None, None),
    ("arm2_N10_N", "arm2", 10, "N", {"expected_reduction_min_SQP_DDP": -100}, None, None),
    ("arm2_N10_SS_xg3", "arm2", 10, "PCG_SS", {"expected_reduction_min_SQP_DDP": -100}, None, [-1.18, -1.58, 0.0, 0.0]),
    ("arm2_N32_SS", "arm2", 32, "PCG_SS", {"expected_reduction_min_SQP_DDP": -100}, None, None),
    ("arm3_N10_SS_default", "arm3", 10, "PCG_SS", {}, None, None),
    ("arm6_N8_SS_default", "arm6", 8, "PCG_SS", {}, None, None),
    ("arm6_N64_SS", "arm6", 64, "PCG_SS", {"expected_reduction_min_SQP_DDP": -100}, None, None),
    ("arm6_N16_SS_semi", "arm6", 16, "PCG_SS", {"expected_reduction_min_SQP_DDP": -100}, None, None),
    ("pend_N20_SS", "pend", 20, "PCG_SS", {"expected_reduction_min_SQP_DDP": -100}, None, None),
    ("pend_N20_SS_qp2", "pend", 20, "PCG_SS", {"expected_reduction_min_SQP_DDP": -100}, {"torque": ([2.0], [-2.0], "QUADRATIC_PENALTY")}, None),
    ("pend_N20_SS_al01", "pend", 20, "PCG_SS", {"expected_reduction_min_SQP_DDP": -100}, {"torque": ([0.1], [-0.1], "AUGMENTED_LAGRANGIAN")}, None),
    ("pend_N20_SS_qp01", "pend", 20, "PCG_SS", {"expected_reduction_min_SQP_DDP": -100}, {"torque": ([0.1], [-0.1], "QUADRATIC_PENALTY")}, None),
    # hard (ACTIVE_SET) torque limits, exact methods ("parity target for N / S only"; known answer 4 SQP, J = 61.41085)
    ("pend_N20_S_as01", "pend", 20, "S", {"expected_reduction_min_SQP_DDP": -100}, {"torque": ([0.1], [-0.1], "ACTIVE_SET")}, None),
    ("pend_N20_N_as01", "pend", 20, "N", {"expected_reduction_min_SQP_DDP": -100}, {"torque": ([0.1], [-0.1], "ACTIVE_SET")}, None),
    ("pend_N20_S_as2", "pend", 20, "S", {"expected_reduction_min_SQP_DDP": -100}, {"torque": ([2.0], [-2.0], "ACTIVE_SET")}, None),
    # integrator types 2 (midpoint, tag suffix _mid) and 3 (rk3, _rk3) exactly as the reference computes them
    ("pend_N20_SS_mid", "pend", 20, "PCG_SS", {"expected_reduction_min_SQP_DDP": -100}, None, None),
    ("pend_N20_SS_rk3", "pend", 20, "PCG_SS", {"expected_reduction_min_SQP_DDP": -100}, None, None),
    ("pend_N20_SS_qp2_mid", "pend", 20, "PCG_SS", {"expected_reduction_min_SQP_DDP": -100}, {"torque": ([2.0], [-2.0], "QUADRATIC_PENALTY")}, None),
    ("arm2_N10_SS_mid", "arm2", 10, "PCG_SS", {"expected_reduction_min_SQP_DDP": -100}, None, None),
    ("arm2_N10_S_rk3", "arm2", 10, "S", {"expected_reduction_min_SQP_DDP": -100}, None, None),
    ("arm2_N10_N_mid", "arm2", 10, "N", {"expected_reduction_min_SQP_DDP": -100}, None, None),
    ("arm6_N16_SS_rk3", "arm6", 16, "PCG_SS", {"expected_reduction_min_SQP_DDP": -100}, None, None),
]


def sec_solve(R):
    """Complete SQP solves from x=0,u=0 with the unmodified reference (SURVEY.md appendix D known answers)."""
    out = {}
    meta = {}
    if os.path.exists(os.path.join(HERE, "solve.npz")) and not os.environ.get("FORCE"):
        out = dict(np.load(os.path.join(HERE, "solve.npz")))
        with open(os.path.join(HERE, "solve_meta.json")) as f:
            meta = json.load(f)
    for tag, name, N, meth, opts, limits, xg in SOLVE_CASES:
        if tag in meta:
            continue
        integ = integ_of(tag)
        plant, cost, cons, solver, xg = make_problem(R, name, N, limits=limits, integrator=integ, xg=xg)
        n = plant.get_num_pos(); nx = 2 * n
        x0 = np.zeros((nx, N)); u0 = np.zeros((n, N - 1))
        o = dict(opts); o["overloading"] = False
        rows = []
        # the reference re-creates self.trace every outer iteration (:555): wrap to keep all rows
        import TrajoptMPCReference as TM
        orig = solver.check_and_update_soft_constraints
        def wrapped(x, u, it, options, _orig=orig, _rows=rows, _solver=solver):
            _rows.extend(_solver.trace)
            return _orig(x, u, it, options)
        solver.check_and_update_soft_constraints = wrapped
        with quiet():
            x, u, e1, e2, outer, it = solver.SQP(x0, u0, N, 0.1, getattr(R.SQPSolverMethods, meth), options=o)
        J = float(solver.totalCost(x, u, N)) if True else 0.0
        with quiet():
            cv = float(solver.totalHardConstraintViolation(x, u, x0[:, 0], N, 0.1))
        out[tag + "/x"] = np.array(x); out[tag + "/u"] = np.array(u); out[tag + "/xg"] = np.asarray(xg, float)
        out[tag + "/exits"] = np.array([e1, e2, outer, it])
        out[tag + "/J"] = np.float64(J); out[tag + "/c"] = np.float64(cv)
        out[tag + "/pcg_iters"] = np.array([len(t[0][0]) - 1 for t in solver.saved_inner_traces], dtype=np.int64)
        tr = [r for r in rows if r["D"] is not None]
        out[tag + "/tr_outer"] = np.array([r["outer_iteration"] for r in tr]); out[tag + "/tr_iter"] = np.array([r["iteration"] for r in tr])
        out[tag + "/tr_ls"] = np.array([r["line_search_iteration"] for r in tr]); out[tag + "/tr_alpha"] = np.array([float(r["alpha"]) for r in tr])
        out[tag + "/tr_rho"] = np.array([float(r["rho"]) for r in tr]); out[tag + "/tr_J"] = np.array([float(r["J"]) for r in tr])
        out[tag + "/tr_c"] = np.array([float(r["c"]) for r in tr]); out[tag + "/tr_ok"] = np.array([bool(r["succeeded_line_search"]) for r in tr])
        if cons is not None:
            out[tag + "/mu"] = cons.torque_limits.quadratic_penalty_mu.copy(); out[tag + "/lam"] = cons.torque_limits.augmented_lagrangian_lambda.copy()
        meta[tag] = dict(robot=name, N=N, method=meth, options=opts, integrator=integ,
                         limits={k: [list(v[0]), list(v[1]), v[2]] for k, v in (limits or {}).items()})
        print(tag, "exits", (e1, e2, outer, it), "J", J, "c", cv, "pcg", out[tag + "/pcg_iters"].tolist(), "ls", (out[tag + "/tr_ls"] + 1).tolist())
    np.savez_compressed(os.path.join(HERE, "solve.npz"), **out)
    with open(os.path.join(HERE, "solve_meta.json"), "w") as f:
        json.dump(meta, f, indent=1)
    print("solve.npz written")


def sec_refdata(R):
    """Re-pack the runs RECORDED BY THE REFERENCE'S AUTHORS (data/3, data/4: arm2, UrdfCost, N=10, PCG-SS) into npz."""
    import pandas as pd
    out = {}
    for run, xg in (("4", [-1.0, 1.5, 0.0, 0.0]), ("3", [-1.18, -1.58, 0.0, 0.0])):
        d = os.path.join(refshim.REF, "data", run)
        ld = lambda f: pd.read_pickle(os.path.join(d, f))
        out[run + "/xg"] = np.array(xg)
        out[run + "/final_x"] = pd.read_csv(os.path.join(d, "final_traj.csv"), index_col=0).values
        out[run + "/final_u"] = pd.read_csv(os.path.join(d, "final_input.csv"), index_col=0).values
        tr = ld("trace.plk").to_dict("records")[1:]          # row 0 is the seed row (D = None)
        for key in ("alpha", "rho", "J", "c", "merit", "D", "reduction_ratio"):
            out["%s/tr_%s" % (run, key)] = np.array([float(r[key]) for r in tr])
        out[run + "/tr_ls"] = np.array([int(r["line_search_iteration"]) for r in tr])
        rows = ld("inner_traces.plk").values.tolist()
        out[run + "/pcg_iters"] = np.array([len(r[0][0]) - 1 for r in rows])
        res = ld("results.plk").values[:, 0]
        out[run + "/exits"] = np.array([int(res[6]), int(res[7]), int(res[8]), int(res[9])])
        out[run + "/J_final"] = np.float64(res[1])
        # every QP solve's dense matrices and every accepted iterate, as recorded by the authors
        for key in ("G", "g", "C", "c", "dxul", "Pinv", "invG", "x", "u"):
            recs = ld(key + ".plk").to_dict("records")
            vals = [np.array(r["value"], dtype=float) for r in recs]
            if key == "invG":      # appended twice per solve in the overloading branch only; keep one per solve
                vals = vals[:len(out[run + "/pcg_iters"])]
            out["%s/%s" % (run, key)] = np.stack(vals)
        # integrator outputs of the first QP solve (x = 0, u = 0): 9 knots
        for key in ("Ak", "Bk", "xkp1"):
            recs = ld(key + ".plk").to_dict("records")
            out["%s/%s_first" % (run, key)] = np.stack([np.array(r["value"], dtype=float).reshape(np.array(recs[0]["value"]).shape) for r in recs[:9]])
        # cost- and plant-level lists (one entry per callback invocation; TrajoptCost.py:411-517, TrajoptPlant.py:297-322): values flattened
        # and NaN-padded to the widest entry, plus the three counters each entry was tagged with
        for key in ("cost", "dx", "grad", "hess", "Jacobian_tot_state", "Minv", "qdd", "dc_du", "dqdd"):
            recs = ld(key + ".plk").to_dict("records")
            vals = [np.asarray(r["value"], dtype=float).reshape(-1) for r in recs]
            w = max(v.size for v in vals)
            arr = np.full((len(vals), w), np.nan)
            for i, v in enumerate(vals):
                arr[i, :v.size] = v
            out["%s/lvl_%s" % (run, key)] = arr
            out["%s/lvl_%s_tags" % (run, key)] = np.array([[int(r["iteration"]), int(r["outer_iteration"]), int(r["line_search_iteration"])] for r in recs])
    np.savez_compressed(os.path.join(HERE, "ref_data.npz"), **out)
    print("ref_data.npz written", {k: v.shape for k, v in out.items() if k.startswith("4/")})


def sec_record_integrators(R):
    """Plant-level lists (URDFPlant.saved_Minv / saved_qdd / saved_dc_du / saved_dqdd, TrajoptPlant.py:297-322) of one complete solve of
    the unmodified reference per integrator type 1, 2, 3 (pend, N = 6, PCG-SS): entry counts, counters and values, i.e. the callback
    sequence of the semi-implicit / midpoint / rk3 branches of TrajoptPlant.integrator."""
    out = {}
    for integ in (1, 2, 3):
        N = 6
        plant, cost, cons, solver, xg = make_problem(R, "pend", N, integrator=integ)
        n = plant.get_num_pos()
        import overloading      # the counters are class attributes that survive a solve (overloading.py:8-10): start each run like a fresh process
        overloading.matrix_.iteration = overloading.matrix_.line_search_iteration = overloading.matrix_.soft_constraint_iteration = 0
        with quiet():
            x, u, e1, e2, outer, it = solver.SQP(np.zeros((2 * n, N)), np.zeros((n, N - 1)), N, 0.1, R.SQPSolverMethods.PCG_SS,
                                                 options={"expected_reduction_min_SQP_DDP": -100, "overloading": False})
        out["%d/xg" % integ] = np.asarray(xg, float); out["%d/x" % integ] = np.array(x); out["%d/u" % integ] = np.array(u)
        out["%d/exits" % integ] = np.array([e1, e2, outer, it])
        for key in ("Minv", "qdd", "dc_du", "dqdd"):
            recs = getattr(plant, "saved_" + key)
            out["%d/lvl_%s" % (integ, key)] = np.stack([np.asarray(r["value"], dtype=float).reshape(-1) for r in recs])
            out["%d/lvl_%s_tags" % (integ, key)] = np.array([[int(r["iteration"]), int(r["outer_iteration"]), int(r["line_search_iteration"])] for r in recs])
        print("integrator", integ, "exits", (e1, e2, outer, it), {k: len(getattr(plant, "saved_" + k)) for k in ("Minv", "qdd", "dc_du", "dqdd")})
    np.savez_compressed(os.path.join(HERE, "record_integrators.npz"), **out)
    print("record_integrators.npz written")


SECTIONS = {"record_integrators": sec_record_integrators, "models": sec_models, "dynamics": sec_dynamics, "integrators": sec_integrators, "kkt": sec_kkt, "solve": sec_solve, "refdata": sec_refdata}

if __name__ == "__main__":
    R = refshim.load()
    which = sys.argv[1:] or list(SECTIONS)
    for s in which:
        SECTIONS[s](R)
