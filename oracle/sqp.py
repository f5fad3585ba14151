"""SQP oracle: outer soft-constraint loop, SQP loop, L1-merit backtracking line search.

Restates /root/reference/TrajoptMPCReference.py: SQP (:510-760), totalCost (:296-310),
totalHardConstraintViolation (:273-294), reduce_regularization (:457-461), check_for_exit_or_error (:463-481),
check_and_update_soft_constraints (:483-508), set_default_options (:91-115).  Control flow per SURVEY.md appendix A.
Trajectories use the reference's layout at the interface: x (nx, N), u (nu, N-1).
"""
import numpy as np

from . import kkt
from . import plant as _plant

METHODS = ("N", "S", "PCG-J", "PCG-BJ", "PCG-SS")


def default_options(options=None):
    o = dict(options or {})
    o.setdefault("exit_tolerance_linSys", 1e-6)
    o.setdefault("max_iter_linSys", 100)
    o.setdefault("exit_tolerance_SQP_DDP", 1e-6)
    o.setdefault("max_iter_SQP_DDP", 100)
    o.setdefault("alpha_factor_SQP_DDP", 0.5)
    o.setdefault("alpha_min_SQP_DDP", 0.005)
    o.setdefault("rho_factor_SQP_DDP", 4)
    o.setdefault("rho_min_SQP_DDP", 1e-3)
    o.setdefault("rho_max_SQP_DDP", 1e3)
    o.setdefault("rho_init_SQP_DDP", 0.001)
    o.setdefault("expected_reduction_min_SQP_DDP", 0.05)
    o.setdefault("expected_reduction_max_SQP_DDP", 3)
    o.setdefault("merit_factor_SQP", 1.5)
    o.setdefault("exit_tolerance_softConstraints", 1e-6)
    o.setdefault("max_iter_softConstraints", 10)
    return o


def total_cost(cost, cons, X, U):
    """totalCost (:296-310): sum of stage costs, then the soft-constraint values knot by knot."""
    J = 0.0
    for v in cost.values(X, U):
        J = J + v
    if cons is not None and cons.any():
        for v in cons.values(X, U):
            J = J + v
    return float(J)


def total_violation(model, X, U, xs, dt, integrator_type=0, gravity=-9.81, cons=None):
    """totalHardConstraintViolation (:273-294), L1 norm of the initial-state and dynamics defects, then of the ACTIVE hard-constraint
    values knot by knot (:285-293)."""
    N = X.shape[0]
    xkp1 = _plant.integrator(model, X[:N - 1], U, dt, integrator_type, False, gravity)
    c = float(np.sum(np.abs(X[0] - xs)))
    for k in range(N - 1):
        c = c + float(np.sum(np.abs(X[k + 1] - xkp1[k])))
    if cons is not None and cons.any_hard():
        for _, vals in cons.hard_rows(X, U):
            if len(vals):
                c = c + float(np.sum(np.abs(vals)))
    return c


def solve_qp(model, cost, cons, X, U, xs, dt, rho, method, o, integrator_type=0, gravity=-9.81, record=None, dense=False):
    """One QP step direction: returns (dz (N,m), l (N,nx), pcg_trace or None).
    dense=True runs the reference's literal dense formulation (oracle.dense) instead of the block form."""
    nx = X.shape[1]
    blocks = kkt.form_blocks(model, cost, cons, X, U, xs, dt, integrator_type, gravity)
    trace = None
    if "hard" in blocks:
        if method not in ("N", "S"):
            raise ValueError("hard (ACTIVE_SET) limits: exact methods N / S only (the reference hands PCG a Schur complement whose size "
                             "no longer matches block_size * Nblocks)")
        dense = True
    if dense and method != "N":
        from . import dense as _dense
        dz, l, trace, d = _dense.solve_qp_dense(blocks, rho, nx, method, o["exit_tolerance_linSys"], o["max_iter_linSys"])
        if record is not None:
            record.setdefault("blocks", []).append(blocks); record.setdefault("dense", []).append(d)
            record.setdefault("l", []).append(l); record.setdefault("dz", []).append(dz)
        return dz, l, trace
    if method == "N":
        from . import dense
        sol = dense.kkt_solve_dense(blocks, rho, nx)[:, 0]
        N, m = blocks["g"].shape
        nz = m * (N - 1) + nx
        dz = np.zeros((N, m))
        dz[:N - 1] = sol[:m * (N - 1)].reshape(N - 1, m)
        dz[N - 1, :nx] = sol[m * (N - 1):nz]
        l = sol[nz:][blocks["dyn_rows"]].reshape(N, nx)
        sch = None
    else:
        sch = kkt.schur(blocks, rho, nx)
        if o.get("_perturb_S") is not None:      # parity-floor experiments only (scripts/parity_floor.py, make_c4_fixture.py): S -> S (1 +/- ulp)
            sch["Sd"], sch["So"] = o["_perturb_S"](sch["Sd"], sch["So"])
        if method == "S":
            l = kkt.bt_solve_dense(sch["Sd"], sch["So"], sch["gamma"])
        else:
            Pd, Po = kkt.preconditioner(sch["Sd"], sch["So"], method[4:])
            l, trace = kkt.pcg(sch["Sd"], sch["So"], sch["gamma"], Pd, Po, o["exit_tolerance_linSys"], o["max_iter_linSys"])
            if record is not None:
                record.setdefault("Pd", []).append(Pd); record.setdefault("Po", []).append(Po)
        dz = kkt.recover(blocks, sch, l, nx)
    if record is not None:
        record.setdefault("blocks", []).append(blocks); record.setdefault("schur", []).append(sch)
        record.setdefault("l", []).append(l); record.setdefault("dz", []).append(dz)
    return dz, l, trace


def sqp(model, cost, cons, x, u, N, dt, method="PCG-SS", options=None, integrator_type=0, gravity=-9.81, record=None, dense=False):
    """TrajoptMPCReference.SQP (:510-760).  Returns dict(x, u, exit_sqp, exit_soft, outer_iter, sqp_iter, trace,
    pcg_iters, alphas) where trace holds every row of every outer iteration (the reference keeps the last outer
    iteration's rows only, :555)."""
    if method not in METHODS:
        raise ValueError("Invalid QP Solver")
    o = default_options(options)
    X = np.array(x, dtype=np.float64).T.copy()       # (N,nx)
    U = np.array(u, dtype=np.float64).T.copy()       # (N-1,nu)
    nx = X.shape[1]
    xs = X[0].copy()
    exit_sqp = 0
    exit_soft = 0
    outer = 0
    trace_all, pcg_iters, ls_trials = [], [], []
    while True:
        rho = o["rho_init_SQP_DDP"]
        drho = 1
        J = total_cost(cost, cons, X, U)
        c = total_violation(model, X, U, xs, dt, integrator_type, gravity, cons)
        mu = 10
        merit = J + mu * c
        trace_all.append(dict(outer_iteration=outer, iteration=0, line_search_iteration=0, alpha=1, rho=rho, J=J, c=c,
                              merit=merit, D=None, reduction_ratio=None, inner_iters=0, succeeded_line_search=False))
        it = 0
        while True:
            dz, l, ptrace = solve_qp(model, cost, cons, X, U, xs, dt, rho, method, o, integrator_type, gravity, record, dense)
            n_inner = (len(ptrace) - 1) if ptrace is not None else 0
            pcg_iters.append(n_inner)
            alpha = 1
            error = False
            ls = 0
            while True:
                Xn = X - alpha * dz[:, :nx]
                Un = U - alpha * dz[:N - 1, nx:]
                J_new = total_cost(cost, cons, Xn, Un)
                c_new = total_violation(model, Xn, Un, xs, dt, integrator_type, gravity, cons)
                grad = cost.gradients(Xn, Un)
                D = 0.0
                sg = cons.gradients(Xn, Un) if (cons is not None and cons.any()) else None
                for k in range(N):
                    w = grad.shape[1] if k < N - 1 else nx
                    D += float(grad[k, :w] @ dz[k, :w])
                    if sg is not None:
                        D += float(sg[k, :w] @ dz[k, :w])
                merit_new = J_new + mu * c_new
                delta_J = J - J_new
                delta_merit = merit - merit_new
                expected = alpha * (D - mu * c_new)
                with np.errstate(divide="ignore", invalid="ignore"):
                    ratio = np.float64(delta_merit) / np.float64(expected)
                if delta_merit >= 0 and ratio >= o["expected_reduction_min_SQP_DDP"] and ratio <= o["expected_reduction_max_SQP_DDP"]:
                    X, U, J, c, merit = Xn, Un, J_new, c_new, merit_new
                    drho = min(drho / o["rho_factor_SQP_DDP"], 1 / o["rho_factor_SQP_DDP"])       # reduce_regularization
                    rho = max(rho * drho, o["rho_min_SQP_DDP"])
                    trace_all.append(dict(outer_iteration=outer, iteration=it, line_search_iteration=ls, alpha=alpha, rho=rho,
                                          J=J, c=c, merit=merit, D=D, reduction_ratio=float(ratio), inner_iters=n_inner,
                                          succeeded_line_search=True))
                    break
                elif alpha > o["alpha_min_SQP_DDP"]:
                    alpha *= o["alpha_factor_SQP_DDP"]
                    ls += 1
                else:
                    error = True
                    trace_all.append(dict(outer_iteration=outer, iteration=it, line_search_iteration=ls, alpha=alpha, rho=rho,
                                          J=J, c=c, merit=merit, D=D, reduction_ratio=float(ratio), inner_iters=n_inner,
                                          succeeded_line_search=False))
                    break
            ls_trials.append(ls + 1)
            # check_for_exit_or_error (:463-481)
            exit_flag = False
            if error:
                drho = max(drho * o["rho_factor_SQP_DDP"], o["rho_factor_SQP_DDP"])
                rho = max(rho * drho, o["rho_min_SQP_DDP"])
                if rho > o["rho_max_SQP_DDP"]:
                    exit_sqp = 2
                    exit_flag = True
            elif delta_J < o["exit_tolerance_SQP_DDP"]:
                exit_sqp = 1
                exit_flag = True
            if it == o["max_iter_SQP_DDP"] - 1:
                exit_sqp = 3
                exit_flag = True
            else:
                it += 1
            if exit_flag:
                break
        # check_and_update_soft_constraints (:483-508)
        exit_flag = False
        max_c = cons.max_value(X, U) if (cons is not None and cons.any()) else 0
        if max_c < o["exit_tolerance_softConstraints"]:
            exit_soft = 1
            exit_flag = True
        if outer == o["max_iter_softConstraints"] - 1:
            exit_soft = 2
            exit_flag = True
        else:
            outer += 1
        if not exit_flag:
            all_mu = cons.update(X, U)
            if all_mu:
                exit_soft = 3
                exit_flag = True
        if exit_flag:
            break
    return dict(x=X.T.copy(), u=U.T.copy(), exit_sqp=exit_sqp, exit_soft=exit_soft, outer_iter=outer, sqp_iter=it,
                trace=trace_all, pcg_iters=pcg_iters, ls_trials=ls_trials, J=J, c=c)
