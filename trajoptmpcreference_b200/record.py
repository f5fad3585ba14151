"""Recording of solver internals in the reference's formats.

The reference fills `saved_*` lists inside SQP (TrajoptMPCReference.py:46-70, appended at :151-265, :309, :322-325, :369-447, :598,
:674-675) and `examples/exampleHelpers.runSolversSQP(record=True)` (:61-159) pickles them to `data/<id>/<name>.plk` next to
`final_traj.csv`, `final_input.csv` and `results.plk`; `analysis/*.ipynb` and `examples/display_final_traj.py` read those files.

Here the numbers come from the GPU: `TrajoptMPCReference.SQP(..., record=True)` installs an iteration hook (include/b2t.h,
b2t_set_iteration_hook) that reads the KKT blocks of every SQP iteration back and `Recorder` lays them out as the reference's dense
matrices:  z = [x_0,u_0,...,x_{N-1}],  G = blkdiag(G_k) + rho I (what solveKKTSystem_Schur stores, :369),  C rows [I], [-A_k,-B_k,I],
c = [x_0-xs; x_{k+1}-f(x_k,u_k)],  invG,  S (block tridiagonal),  gamma,  Pinv (J / BJ: block diagonal, SS: PCG.py:113-212),  l,
dxul = [dz; l].  Every entry is tagged {'iteration','outer_iteration','line_search_iteration'} like the reference's.
Lists the batched solver has no per-call equivalent of (cost- and plant-level `saved_*`: one entry per Python callback invocation)
are not produced; see INTEGRATION.md.
"""
import os
import time

import numpy as np

from . import _lib

SQP_VARS = ["saved_Pinv", "saved_inner_traces", "saved_G", "trace", "saved_invG", "saved_c", "saved_g", "saved_C", "saved_tot_cost",
            "saved_J_tot_constraints", "saved_Ak", "saved_Bk", "saved_xkp1", "saved_dxul", "saved_x", "saved_u", "saved_S", "saved_gamma",
            "saved_l"]


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
        elif event == _lib.HOOK_STEP:
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
    return results_all


def runSQPExample(plant, cost, constraints, N, dt, solver_methods, options=None, n_test=0, record=False, data_dir="../data"):
    """examples/exampleHelpers.runSQPExample (:161-170)."""
    from .api import TrajoptMPCReference
    solver = TrajoptMPCReference(plant, cost) if constraints is None else TrajoptMPCReference(plant, cost, constraints)
    return runSolversSQP(solver, N, dt, solver_methods, options, n_test, record, data_dir)
