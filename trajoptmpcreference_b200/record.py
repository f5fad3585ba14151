"""Recording of solver internals in the reference's formats.

The reference fills `saved_*` lists inside SQP (TrajoptMPCReference.py:46-70, appended at :151-265, :309, :322-325, :369-447, :598,
:674-675) and `examples/exampleHelpers.runSolversSQP(record=True)` (:61-159) pickles them to `data/<id>/<name>.plk` next to
`final_traj.csv`, `final_input.csv` and `results.plk`; `analysis/*.ipynb` and `examples/display_final_traj.py` read those files.

Here the numbers come from the GPU: `TrajoptMPCReference.SQP(..., record=True)` installs an iteration hook (include/b2t.h,
b2t_set_iteration_hook) that reads the KKT blocks of every SQP iteration back and `Recorder` lays them out as the reference's dense
matrices:  z = [x_0,u_0,...,x_{N-1}],  G = blkdiag(G_k) + rho I (what solveKKTSystem_Schur stores, :369),  C rows [I], [-A_k,-B_k,I],
c = [x_0-xs; x_{k+1}-f(x_k,u_k)],  invG,  S (block tridiagonal),  gamma,  Pinv (J / BJ: block diagonal, SS: PCG.py:113-212),  l,
dxul = [dz; l].  Every entry is tagged {'iteration','outer_iteration','line_search_iteration'} like the reference's.
The cost- and plant-level lists (`UrdfCost.saved_cost / saved_dx / saved_grad / saved_hess / saved_Jacobian_tot_state`,
TrajoptCost.py:411-421, 457-458, 516-517; `URDFPlant.saved_c / saved_Minv / saved_qdd / saved_dc_du / saved_dqdd`, TrajoptPlant.py:297-299,
318-322) hold one entry per Python callback invocation in the reference.  The batched solver makes no such calls, so the Recorder
REPLAYS the reference's call sequence from the recorded iterates: per outer iteration totalCost (:296-310) and
totalHardConstraintViolation (:273-294) of the start point; per QP solve the loop of formKKTSystemBlocks (:200-271: hessian, gradient,
integrator with and without gradient per knot); per line-search trial totalCost, the violation and the directional-derivative loop
(:617-648) at x - alpha dz.  Every knot of a trajectory is evaluated in one launch of the cost / plant kernels
(B2T_ARR_COST_*, B2T_ARR_PLANT_TERMS) and the entries carry the reference's three counters as it would hold them at that call
(`matrix_.iteration`, `.soft_constraint_iteration`, `.line_search_iteration`, incl. their stale values between loops).
"""
import os
import time

import numpy as np

from . import _lib

SQP_VARS = ["saved_Pinv", "saved_inner_traces", "saved_G", "trace", "saved_invG", "saved_c", "saved_g", "saved_C", "saved_tot_cost",
            "saved_J_tot_constraints", "saved_Ak", "saved_Bk", "saved_xkp1", "saved_dxul", "saved_x", "saved_u", "saved_S", "saved_gamma",
            "saved_l"]


COST_VARS = ["saved_cost", "saved_grad", "saved_hess", "saved_Jacobian_tot_state", "saved_dx"]
PLANT_VARS = ["saved_Minv", "saved_c", "saved_dc_du", "saved_qdd", "saved_dqdd"]


class Recorder:
    """Iteration hook of one single-instance solve; fills `owner.saved_*`."""

    def __init__(self, owner, solver, method, xs, options):
        self.o, self.s, self.method = owner, solver, method
        self.xs = np.array(xs, dtype=np.float64).reshape(-1)
        self.tol = float(options.get("exit_tolerance_linSys", 1e-6))
        self.max_iter = int(options.get("max_iter_linSys", 100))
        for name in SQP_VARS:
            if name != "trace":
                setattr(owner, name, [])
        self._tag = None
        self._last_xu = None
        # replay of the cost- / plant-level callbacks (see the module docstring)
        self.alpha_factor = float(options.get("alpha_factor_SQP_DDP", 0.5))
        self.max_iter_sqp = int(options.get("max_iter_SQP_DDP", 100))
        cost, plant = getattr(owner, "cost", None), getattr(owner, "plant", None)
        self._replay = plant is not None and hasattr(plant, "saved_c")      # a real TrajoptMPCReference (host-side layout tests use stand-ins)
        self._cost_lists = hasattr(cost, "saved_cost")
        self._ref_it, self._ref_ls, self._outer_seen = 0, 0, -1
        self._probe = None
        self._pre = None                    # (X, U, dz, iteration, outer, total_trials) of the QP solve whose line search is running
        for obj, names in ((cost, ("cost", "grad", "hess", "Jacobian_tot_state", "dx")), (plant, ("c", "Minv", "qdd", "dc_du", "dqdd"))):
            for name in names:
                if hasattr(obj, "saved_" + name):
                    setattr(obj, "saved_" + name, [])
            if hasattr(obj, "recording"):
                obj.recording = True        # later direct calls (e.g. runSolversSQP's post-processing) append like the reference's do

    # ---- cost / plant kernels over all knots of one trajectory
    def _eval(self, X, U):
        from .api import BatchSolver, split_plant_terms
        s = self.s
        itype = self.o.plant.integrator_type
        if self._probe is None:       # plant terms and dqdd at given points do not depend on the integrator: the Euler kernels serve all types
            self._probe = BatchSolver(self.o.plant, self.o.cost, None, N=s.N, dt=s.dt, batch=1, integrator_type=itype if itype in (0, 1) else 0)
        p = self._probe
        p.set_goals(np.asarray(self.o.cost.xg, dtype=np.float64).reshape(1, -1))
        p.set_trajectory(X[None], U[None])
        n, N = s.n, s.N
        out = dict(value=p.fetch("cost_value")[0, :, 0], dx=p.fetch("cost_err")[0], grad=p.fetch("cost_grad")[0],
                   hess=p.fetch("cost_hess")[0].reshape(N, s.m, s.m), jtot=p.fetch("cost_jtot")[0].reshape(N, s.nx, s.nx))

        def plant_at(P):
            p.set_trajectory(P[None], U[None])
            p.stage_dynamics()
            return [split_plant_terms(t, n) for t in p.fetch("plant_terms")[0]], p.fetch("dqdd")[0].reshape(N, n, 3 * n)
        # the stage points of the plant's integrator (TrajoptPlant.py:140-175): point_s = x_k + h_s dt [qd_k ; qdd(point_{s-1}, u_k)]
        stages = [plant_at(X)]
        for h in {2: (0.5,), 3: (0.5, 0.75)}.get(itype, ()):
            qdd = np.stack([t[1] for t in stages[-1][0]], axis=1)           # (n, N)
            stages.append(plant_at(X + h * s.dt * np.vstack([X[n:], qdd])))
        out["stages"] = stages
        return out

    def _put(self, obj, tag, **entries):
        for name, value in entries.items():
            lst = getattr(obj, "saved_" + name, None)
            if lst is not None:
                lst.append(dict(value=value, iteration=tag[0], outer_iteration=tag[1], line_search_iteration=tag[2]))

    def _integrator(self, e, k, tag, gradient=False):
        """The plant callbacks of ONE plant.integrator(x_k, u_k, ...) call: forward_dynamics at every stage point (TrajoptPlant.py:95,
        :118, :141-143, :171-175; appends :297-299), then -- with return_gradient -- forward_dynamics_gradient at every stage point
        (:100, :131, :150-156, :185-198; appends :318-322).  The semi-implicit branch calls the gradient without the counters (:131)."""
        for terms, _ in e["stages"]:
            c, qdd, Minv, dc_du = terms[k]
            self._put(self.o.plant, tag, c=c, Minv=Minv, qdd=qdd)
        if gradient:
            gtag = (0, 0, 0) if self.o.plant.integrator_type == 1 else tag
            for terms, dqdd in e["stages"]:
                c, qdd, Minv, dc_du = terms[k]
                self._put(self.o.plant, gtag, Minv=Minv, c=c, qdd=qdd, dc_du=dc_du, dqdd=dqdd[k].copy())

    def _cost(self, e, k, tag, what):
        if not self._cost_lists:
            return
        s, ne = self.s, 4
        w = s.m if k < s.N - 1 else s.nx
        jt = e["jtot"][k][:ne].copy()
        if what == "value":
            self._put(self.o.cost, tag, cost=float(e["value"][k]), dx=e["dx"][k][:ne].copy())
        elif what == "grad":
            self._put(self.o.cost, tag, grad=e["grad"][k][:w].copy(), Jacobian_tot_state=jt)
        else:
            self._put(self.o.cost, tag, hess=e["hess"][k][:w, :w].copy(), Jacobian_tot_state=jt)

    def _total_cost_and_violation(self, e, tag):
        N = self.s.N
        for k in range(N):                    # totalCost (:296-310)
            self._cost(e, k, tag, "value")
        for k in range(N - 1):                # totalHardConstraintViolation (:273-294): integrator -> forward_dynamics
            self._integrator(e, k, tag)

    def _replay_kkt(self, e, tag):
        N = self.s.N
        ptag = (tag[0], tag[1], 0)            # the numpy branch passes no iter_3 to plant.integrator here (:227, :230): default 0
        for k in range(N - 1):                # formKKTSystemBlocks (:200-271)
            self._cost(e, k, tag, "hess")
            self._cost(e, k, tag, "grad")
            self._integrator(e, k, ptag, gradient=True)      # Ak, Bk (:227)
            self._integrator(e, k, ptag)                     # x_{k+1} (:230)
        self._cost(e, N - 1, tag, "hess")
        self._cost(e, N - 1, tag, "grad")

    def _replay_line_search(self, trials):
        X, U, dz, it, outer = self._pre
        s = self.s
        for j in range(trials):               # SQP :606-744: one trial per alpha = alpha_factor^j
            a = self.alpha_factor ** j
            Xn = X - a * dz[:, :s.nx].T
            Un = U - a * dz[:s.N - 1, s.nx:].T
            e = self._eval(Xn, Un)
            tag = (it, outer, j)
            self._total_cost_and_violation(e, tag)
            for k in range(s.N):              # directional derivative D (:635-648)
                self._cost(e, k, tag, "grad")
        self._ref_ls = trials - 1

    # ---- dense layouts of the reference
    def _dense(self):
        s = self.s
        N, nx, nu, m = s.N, s.nx, s.nu, s.m
        nz = m * (N - 1) + nx
        H = s.fetch("kkt_hess")[0].reshape(N, m, m)
        gk = s.fetch("g")[0]
        AB = s.fetch("AB")[0].reshape(N, nx, m)
        X = s.fetch("x")[0]
        xkp1 = s.fetch("xkp1")[0]
        Gh = s.fetch("Ghat")[0].reshape(N, m, m)
        Sd = s.fetch("Sd")[0].reshape(N, nx, nx)
        So = s.fetch("So")[0].reshape(N, nx, nx)
        Pd = s.fetch("Pd")[0].reshape(N, nx, nx)
        gam = s.fetch("gamma")[0]
        l = s.fetch("l")[0]
        dz = s.fetch("dz")[0]
        rho = float(s.get_scalars()[0, 3])
        G = np.zeros((nz, nz)); invG = np.zeros((nz, nz)); g = np.zeros((nz, 1)); dxu = np.zeros((nz, 1))
        C = np.zeros((nx * N, nz)); c = np.zeros((nx * N, 1))
        S = np.zeros((nx * N, nx * N)); Pinv = np.zeros((nx * N, nx * N))
        C[:nx, :nx] = np.eye(nx)
        c[:nx, 0] = X[0] - self.xs
        for k in range(N):
            w = m if k < N - 1 else nx
            a = k * m
            G[a:a + w, a:a + w] = H[k, :w, :w] + rho * np.eye(w)
            invG[a:a + w, a:a + w] = Gh[k, :w, :w]
            g[a:a + w, 0] = gk[k, :w]
            dxu[a:a + w, 0] = dz[k, :w]
            r = k * nx
            S[r:r + nx, r:r + nx] = Sd[k]
            Pinv[r:r + nx, r:r + nx] = Pd[k]
            if k < N - 1:
                C[r + nx:r + 2 * nx, a:a + m] = -AB[k]
                C[r + nx:r + 2 * nx, a + m:a + m + nx] = np.eye(nx)
                c[r + nx:r + 2 * nx, 0] = X[k + 1] - xkp1[k]
            if k > 0:
                S[r:r + nx, r - nx:r] = So[k]
                S[r - nx:r, r:r + nx] = So[k].T
        if self.method.name == "PCG_SS":      # Pinv = D^-1 - D^-1 O D^-1 restricted to the block tridiagonal (PCG.py:181-211)
            for k in range(1, N):
                r = k * nx
                blk = -Pd[k] @ So[k] @ Pd[k - 1]
                Pinv[r:r + nx, r - nx:r] = blk
                Pinv[r - nx:r, r:r + nx] = blk.T
        return dict(G=G, g=g, C=C, c=c, invG=invG, S=S, gamma=gam.reshape(-1, 1), Pinv=Pinv, l=l.reshape(-1, 1),
                    dxul=np.vstack([dxu, l.reshape(-1, 1)]), A=AB[:N - 1, :, :nx], B=AB[:N - 1, :, nx:], xkp1=xkp1[:N - 1],
                    gck=gk - s.fetch("cost_grad")[0])

    def __call__(self, event, ipass):
        s, o = self.s, self.o
        st = s.get_status()[0]
        if event == _lib.HOOK_LINSYS:
            self._tag = {"iteration": int(st[3]), "outer_iteration": int(st[2]), "line_search_iteration": 0}
            if self._last_xu is None:
                self._last_xu = s.get_trajectory()
            d = self._dense()
            tag = self._tag
            # cost- / plant-level callbacks the reference makes up to and inside this QP solve
            it, outer = int(st[3]), int(st[2])
            if self._replay:
                Xc, Uc = s.get_trajectory()
                e = self._eval(Xc[0], Uc[0])
                if outer != self._outer_seen:     # start of an outer iteration: J and c of the start point, counters still stale (:541-542)
                    self._outer_seen = outer
                    self._total_cost_and_violation(e, (self._ref_it, outer, self._ref_ls))
                self._ref_it = it
                self._replay_kkt(e, (it, outer, self._ref_ls))

            def put(name, value):
                getattr(o, name).append(dict(value=value, **tag))
            for k in range(s.N - 1):
                put("saved_Ak", d["A"][k]); put("saved_Bk", d["B"][k]); put("saved_xkp1", d["xkp1"][k])
                if s.has_limits:
                    put("saved_J_tot_constraints", d["gck"][k].reshape(-1, 1))
            put("saved_G", d["G"]); put("saved_g", d["g"]); put("saved_C", d["C"]); put("saved_c", d["c"])
            put("saved_invG", d["invG"]); put("saved_S", d["S"]); put("saved_gamma", d["gamma"])
            if self.method.name.startswith("PCG"):
                nu = s.fetch("nu_trace")[0]
                n_it = next((i for i in range(1, min(self.max_iter, 127) + 1) if nu[i] < self.tol), min(self.max_iter, 127))
                o.saved_inner_traces.append(((nu[:n_it + 1].tolist(), []), tag["iteration"], tag["outer_iteration"]))
                put("saved_Pinv", d["Pinv"])
            put("saved_l", d["l"]); put("saved_dxul", d["dxul"])
            if self._replay:
                self._pre = (Xc[0].copy(), Uc[0].copy(), s.fetch("dz")[0].copy(), it, outer)
                self._trials_before = int(st[6])
        elif event == _lib.HOOK_STEP:
            if self._pre is not None:
                self._replay_line_search(int(st[6]) - self._trials_before)
                it = self._pre[3]
                self._ref_it = it if it == self.max_iter_sqp - 1 else it + 1          # check_for_exit_or_error (:476-479)
                self._pre = None
            x, u = s.get_trajectory()
            if self._last_xu is None or not (np.array_equal(x, self._last_xu[0]) and np.array_equal(u, self._last_xu[1])):
                o.saved_x.append(dict(value=x[0].copy(), **self._tag))
                o.saved_u.append(dict(value=u[0].copy(), **self._tag))
                o.saved_tot_cost.append(dict(value=float(s.get_scalars()[0, 0]), **self._tag))
            self._last_xu = (x, u)


def save_in_file(file_path, value, csv=False):
    """examples/exampleHelpers.py:41-55: pandas DataFrame, pickled (.plk) or CSV; a 'saved_' prefix is dropped from the file name."""
    import pandas as pd
    if "saved_" in file_path:
        file_path = file_path.replace("saved_", "")
    directory = os.path.dirname(file_path)
    if directory and not os.path.exists(directory):
        os.makedirs(directory)
    df = pd.DataFrame(value)
    if csv:
        df.to_csv(file_path)
    else:
        df.to_pickle(file_path)


def runSolversSQP(trajoptMPCReference, N, dt, solver_methods, options=None, n_test=0, record=False, data_dir="../data"):
    """examples/exampleHelpers.runSolversSQP (:61-159): zero initial guess, one SQP call per method; with record=True writes
    `<data_dir>/<n_test>/{final_traj.csv, final_input.csv, results.plk, <var>.plk}`.  Returns the list of SQP result tuples."""
    from .api import QuadraticCost, UrdfCost
    options = {} if options is None else options
    results_all = []
    for solver in solver_methods:
        nq = trajoptMPCReference.plant.get_num_pos()
        nv = trajoptMPCReference.plant.get_num_vel()
        nx = nq + nv
        nu = trajoptMPCReference.plant.get_num_cntrl()
        x = np.zeros((nx, N)); u = np.zeros((nu, N - 1))
        t1 = time.time()
        x, u, exit_sqp, exit_soft, outer_iter, sqp_iter = trajoptMPCReference.SQP(x, u, N, dt, LINEAR_SYSTEM_SOLVER_METHOD=solver, options=options,
                                                                                  record=record)
        t2 = time.time()
        results_all.append((x, u, exit_sqp, exit_soft, outer_iter, sqp_iter))
        if not record:
            continue
        base = os.path.join(data_dir, str(n_test))
        save_in_file(os.path.join(base, "final_traj.csv"), x, csv=True)
        save_in_file(os.path.join(base, "final_input.csv"), u, csv=True)
        cost = trajoptMPCReference.cost
        J = Jx = Ju = 0.0
        for k in range(N - 1):
            full, state = cost.value(x[:, k], u[:, k]), cost.value(x[:, k], None)
            J += full; Jx += state; Ju += full - state
        last = cost.value(x[:, N - 1], None)
        J += last; Jx += last
        E = float(np.dot(u[:, -1], x[nq:, -1])) * dt * 10000
        error = 0
        if isinstance(cost, UrdfCost):
            error = cost.delta_x(x[:, -1])
        elif isinstance(cost, QuadraticCost):
            error = x[:, -1] - cost.xg
        save_in_file(os.path.join(base, "results.plk"), [t2 - t1, J, Jx, Ju, error, E, exit_sqp, exit_soft, outer_iter, sqp_iter])
        for var in SQP_VARS:
            save_in_file(os.path.join(base, var + ".plk"), getattr(trajoptMPCReference, var))
        # plant- and cost-level lists (exampleHelpers.py:140-154; the plant's saved_c is written after, and over, the solver's c.plk
        # there -- kept apart here as plant_c.plk, with c.plk holding the solver-level list the analysis notebooks read)
        for var in PLANT_VARS:
            if hasattr(trajoptMPCReference.plant, var):
                save_in_file(os.path.join(base, ("plant_c" if var == "saved_c" else var) + ".plk"), getattr(trajoptMPCReference.plant, var))
        for var in COST_VARS:
            if hasattr(cost, var):
                save_in_file(os.path.join(base, var + ".plk"), getattr(cost, var))
    return results_all


def runSQPExample(plant, cost, constraints, N, dt, solver_methods, options=None, n_test=0, record=False, data_dir="../data"):
    """examples/exampleHelpers.runSQPExample (:161-170)."""
    from .api import TrajoptMPCReference
    solver = TrajoptMPCReference(plant, cost) if constraints is None else TrajoptMPCReference(plant, cost, constraints)
    return runSolversSQP(solver, N, dt, solver_methods, options, n_test, record, data_dir)
