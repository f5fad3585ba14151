"""Dense re-assembly of the reference's matrices from the oracle's blocks (cross-check only).

Follows TrajoptMPCReference.formKKTSystemBlocks (:200-271) / solveKKTSystem_Schur (:415-452) literally: dense G,
C, invG = np.linalg.inv(G), S = -C invG C^T, gamma = c - C invG g, and PCG.compute_preconditioner's dense Pinv."""
import numpy as np


def assemble_bt(Sd, So):
    N, nx, _ = Sd.shape
    S = np.zeros((N * nx, N * nx))
    for k in range(N):
        S[k * nx:(k + 1) * nx, k * nx:(k + 1) * nx] = Sd[k]
        if k > 0:
            S[k * nx:(k + 1) * nx, (k - 1) * nx:k * nx] = So[k - 1]
            S[(k - 1) * nx:k * nx, k * nx:(k + 1) * nx] = So[k - 1].T
    return S


def assemble_kkt(blocks, nx):
    """Dense G, g, C, c in the reference's row order (:200-271): initial-state rows, then per knot the dynamics rows of knot k+1
    followed by the ACTIVE hard-constraint rows of knot k (:238-248), finally the hard rows of the terminal knot (:263-270).
    blocks["dyn_rows"] receives the indices of the nx N dynamics / initial-state rows inside C."""
    G, g, A, B, c = blocks["G"], blocks["g"], blocks["A"], blocks["B"], blocks["c"]
    N, m, _ = G.shape
    nz = m * (N - 1) + nx
    hard = blocks.get("hard")
    nh = sum(len(v) for _, v in hard) if hard is not None else 0
    Gd = np.zeros((nz, nz)); gd = np.zeros((nz, 1))
    C = np.zeros((nx * N + nh, nz)); cd = np.zeros((nx * N + nh, 1))
    C[:nx, :nx] = np.eye(nx)
    cd[:nx, 0] = c[0]
    dyn = list(range(nx))
    r = nx
    for k in range(N - 1):
        Gd[k * m:(k + 1) * m, k * m:(k + 1) * m] = G[k]
        gd[k * m:(k + 1) * m, 0] = g[k]
        C[r:r + nx, k * m:k * m + m + nx] = np.hstack((-A[k], -B[k], np.eye(nx)))
        cd[r:r + nx, 0] = c[k + 1]
        dyn += list(range(r, r + nx))
        r += nx
        if hard is not None and len(hard[k][1]):
            rows, vals = hard[k]
            C[r:r + len(vals), k * m:(k + 1) * m] = rows
            cd[r:r + len(vals), 0] = vals
            r += len(vals)
    Gd[(N - 1) * m:, (N - 1) * m:] = G[N - 1, :nx, :nx]
    gd[(N - 1) * m:, 0] = g[N - 1, :nx]
    if hard is not None and len(hard[N - 1][1]):
        rows, vals = hard[N - 1]
        C[r:r + len(vals), (N - 1) * m:] = rows[:, :nx]
        cd[r:r + len(vals), 0] = vals
    blocks["dyn_rows"] = np.array(dyn)
    return Gd, gd, C, cd


def schur_dense(blocks, rho, nx):
    Gd, gd, C, cd = assemble_kkt(blocks, nx)
    Gd = Gd + rho * np.eye(Gd.shape[0])
    invG = np.linalg.inv(Gd)
    S = -np.matmul(C, np.matmul(invG, C.T))
    gamma = cd - np.matmul(C, np.matmul(invG, gd))
    return dict(G=Gd, g=gd, C=C, c=cd, invG=invG, S=S, gamma=gamma)


def kkt_solve_dense(blocks, rho, nx):
    """Method 'N' (solveKKTSystem :313-359)."""
    Gd, gd, C, cd = assemble_kkt(blocks, nx)
    Gd = Gd + rho * np.eye(Gd.shape[0]) if rho != 0 else Gd
    nc = C.shape[0]
    KKT = np.hstack((np.vstack((Gd, C)), np.vstack((C.T, np.zeros((nc, nc))))))
    rhs = np.vstack((gd, cd))
    return np.linalg.solve(KKT, rhs)


def preconditioner_dense(S, nx, kind):
    """PCG.compute_preconditioner numpy branch (PCG.py:166-212), literal, on the dense S."""
    if kind == "0":
        return np.identity(S.shape[0])
    if kind == "J":
        return np.linalg.inv(np.diag(np.diag(S)))
    nb = int(S.shape[0] / nx)
    P = np.zeros(S.shape)
    sl = lambda k: slice(k * nx, (k + 1) * nx)
    if kind == "BJ":
        for k in range(nb):
            P[sl(k), sl(k)] = np.linalg.inv(S[sl(k), sl(k)])
        return P
    for k in range(nb):
        P[sl(k), sl(k)] = np.linalg.inv(S[sl(k), sl(k)])
        if k % 2:
            P[sl(k), sl(k - 1)] = -np.matmul(P[sl(k), sl(k)], np.matmul(S[sl(k), sl(k - 1)], P[sl(k - 1), sl(k - 1)]))
        elif k > 0:
            P[sl(k - 1), sl(k)] = -np.matmul(P[sl(k - 1), sl(k - 1)], np.matmul(S[sl(k - 1), sl(k)], P[sl(k), sl(k)]))
    for k in range(nb):
        if k % 2:
            P[sl(k - 1), sl(k)] = P[sl(k), sl(k - 1)].transpose()
            if k < nb - 1:
                P[sl(k + 1), sl(k)] = P[sl(k), sl(k + 1)].transpose()
    return P


def pcg_dense(A, b, Pinv, tol=1e-6, max_iter=100):
    """PCG.pcg numpy branch (PCG.py:66-111), literal (zero initial guess)."""
    x = np.zeros((A.shape[0], 1))
    r = b - (A @ x)
    r_tilde = Pinv @ r
    p = r_tilde
    nu = r.transpose() @ r_tilde
    trace = nu[0].tolist()
    for _ in range(max_iter):
        Ap = A @ p
        alpha = nu / (p.transpose() @ Ap)
        r = r - Ap * alpha
        x = x + p * alpha
        r_tilde = Pinv @ r
        nu_prime = r.transpose() @ r_tilde
        trace.append(nu_prime.tolist()[0][0])
        if abs(nu_prime) < tol:
            break
        beta = nu_prime / nu
        p = r_tilde + p * beta
        nu = nu_prime
    return x, list(map(abs, trace))


def solve_qp_dense(blocks, rho, nx, method, tol=1e-6, max_iter=100):
    """solveKKTSystem_Schur numpy branch (:415-452), literal.  Returns (dz (N,m), l (N,nx), pcg trace)."""
    d = schur_dense(blocks, rho, nx)
    N, m = blocks["g"].shape
    trace = None
    if method == "S":
        l = np.linalg.solve(d["S"], d["gamma"])
    else:
        Pinv = preconditioner_dense(d["S"], nx, method[4:])
        l, trace = pcg_dense(d["S"], d["gamma"], Pinv, tol, max_iter)
        d["Pinv"] = Pinv
    gCl = d["g"] - np.matmul(d["C"].T, l)
    dxu = np.matmul(d["invG"], gCl)[:, 0]
    dz = np.zeros((N, m))
    dz[:N - 1] = dxu[:m * (N - 1)].reshape(N - 1, m)
    dz[N - 1, :nx] = dxu[m * (N - 1):]
    d["l"] = l
    return dz, l[blocks["dyn_rows"], 0].reshape(N, nx), trace, d
