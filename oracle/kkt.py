"""KKT blocks, block-tridiagonal Schur complement, preconditioners, PCG and step recovery.

Block-structured restatement of /root/reference/TrajoptMPCReference.py: formKKTSystemBlocks (:118-271),
solveKKTSystem_Schur (:361-455) and /root/reference/GBD-PCG-Python/PCG.py (compute_preconditioner :113-212,
pcg :66-111).  `oracle.dense` rebuilds the reference's dense matrices from these blocks for cross-checks.

With  Ghat_k = (G_k + rho I)^-1,  AB_k = [A_k B_k]  (SURVEY.md section 8a):
    S_00        = -Ghat_0[:nx,:nx]
    S_k+1,k+1   = -(AB_k Ghat_k AB_k^T + Ghat_k+1[:nx,:nx])
    S_k+1,k     =  AB_k Ghat_k[:, :nx]          (= S_k,k+1^T)
    gamma_0     = c_0 - (Ghat_0 g_0)[:nx]
    gamma_k+1   = c_k+1 + AB_k (Ghat_k g_k) - (Ghat_k+1 g_k+1)[:nx]
    dxu_k       = Ghat_k (g_k - [l_k;0] + AB_k^T l_k+1)
"""
import numpy as np

from . import plant as _plant


def form_blocks(model, cost, cons, X, U, xs, dt, integrator_type=0, gravity=-9.81):
    """formKKTSystemBlocks (:200-271) per knot.  Returns dict with
    G (N,m,m), g (N,m), A (N-1,nx,nx), B (N-1,nx,nu), c (N,nx), xkp1 (N-1,nx)."""
    N, nx = X.shape
    G = cost.hessians(X, U)
    g = cost.gradients(X, U)
    if cons is not None and cons.any():
        gck = cons.gradients(X, U)
        g = g + gck
        G = G + gck[:, :, None] * gck[:, None, :]
    A, B = _plant.integrator(model, X[:N - 1], U, dt, integrator_type, True, gravity)
    xkp1 = _plant.integrator(model, X[:N - 1], U, dt, integrator_type, False, gravity)
    c = np.zeros((N, nx))
    c[0] = X[0] - xs
    c[1:] = X[1:] - xkp1
    blocks = dict(G=G, g=g, A=A, B=B, c=c, xkp1=xkp1)
    if cons is not None and cons.any_hard():
        blocks["hard"] = cons.hard_rows(X, U)       # extra KKT rows (:238-248, :263-270); only the dense assembly handles them
    return blocks


def schur(blocks, rho, nx):
    """Block form of  S = -C inv(G + rho I) C^T,  gamma = c - C inv(G) g   (:419-424).
    Returns Ghat (N,m,m; terminal block embedded top-left), Sd (N,nx,nx), So (N-1,nx,nx) with So[k] = S_{k+1,k}, gamma (N,nx)."""
    G, g, A, B, c = blocks["G"], blocks["g"], blocks["A"], blocks["B"], blocks["c"]
    N, m, _ = G.shape
    Ghat = np.zeros_like(G)
    Gr = G + rho * np.eye(m)
    Ghat[:N - 1] = np.linalg.inv(Gr[:N - 1])
    Ghat[N - 1, :nx, :nx] = np.linalg.inv(Gr[N - 1, :nx, :nx])
    AB = np.concatenate([A, B], axis=-1)                              # (N-1,nx,m)
    W = np.matmul(Ghat[:N - 1], np.swapaxes(AB, -1, -2))              # Ghat AB^T  (m,nx)
    Sd = np.zeros((N, nx, nx))
    Sd[0] = -Ghat[0, :nx, :nx]
    Sd[1:] = -(np.matmul(AB, W) + Ghat[1:, :nx, :nx])
    So = np.matmul(AB, Ghat[:N - 1, :, :nx])                          # S_{k+1,k}
    Gg = np.matmul(Ghat, g[..., None])[..., 0]                         # Ghat g
    gamma = np.zeros((N, nx))
    gamma[0] = c[0] - Gg[0, :nx]
    gamma[1:] = c[1:] + np.matmul(AB, Gg[:N - 1, :, None])[..., 0] - Gg[1:, :nx]
    return dict(Ghat=Ghat, Sd=Sd, So=So, gamma=gamma, AB=AB)


def bt_matvec(Sd, So, p):
    """y = S p for block-tridiagonal symmetric S; p (N,nx)."""
    y = np.matmul(Sd, p[..., None])[..., 0]
    y[1:] += np.matmul(So, p[:-1, :, None])[..., 0]
    y[:-1] += np.matmul(np.swapaxes(So, -1, -2), p[1:, :, None])[..., 0]
    return y


def preconditioner(Sd, So, kind):
    """PCG.compute_preconditioner (PCG.py:166-212).  Returns (Pd, Po): block-tridiagonal Pinv with Po[k] = Pinv_{k+1,k}.

    'SS' (verified against the reference's loops):  Pinv_kk = S_kk^-1,
    Pinv_{k,k-1} = -S_kk^-1 S_{k,k-1} S_{k-1,k-1}^-1 computed as -(Pinv_kk (S_{k,k-1} Pinv_{k-1,k-1})) for odd k and as
    the transpose of -(Pinv_{k-1,k-1} (S_{k-1,k} Pinv_kk)) for even k; the upper blocks are exact transposes."""
    N, nx, _ = Sd.shape
    if kind == "0":
        return np.broadcast_to(np.eye(nx), (N, nx, nx)).copy(), np.zeros((N - 1, nx, nx))
    if kind == "J":
        d = np.einsum("kii->ki", Sd)
        Pd = np.zeros_like(Sd)
        idx = np.arange(nx)
        Pd[:, idx, idx] = 1.0 / d
        return Pd, np.zeros((N - 1, nx, nx))
    Pd = np.linalg.inv(Sd)
    Po = np.zeros((N - 1, nx, nx))
    if kind == "BJ":
        return Pd, Po
    if kind != "SS":
        raise ValueError("Invalid preconditioner options are [0: none, J : Jacobi, BJ: Block-Jacobi, SS: Symmetric Stair]")
    for k in range(1, N):
        if k % 2 == 1:      # odd block row: left-of-diagonal term (PCG.py:190-195)
            Po[k - 1] = -np.matmul(Pd[k], np.matmul(So[k - 1], Pd[k - 1]))
        else:               # even block: right-of-diagonal term of the previous (odd) row, mirrored (:196-210)
            up = -np.matmul(Pd[k - 1], np.matmul(So[k - 1].T, Pd[k]))
            Po[k - 1] = up.T
    return Pd, Po


def pcg(Sd, So, gamma, Pd, Po, tol=1e-6, max_iter=100, guess=None):
    """PCG.pcg (PCG.py:66-111) on the block-tridiagonal system.  Returns (l (N,nx), trace |nu| list).
    iterations = len(trace) - 1."""
    b = gamma
    x = np.zeros_like(b) if guess is None else np.array(guess, dtype=np.float64).reshape(b.shape)
    r = b - bt_matvec(Sd, So, x)
    rt = bt_matvec(Pd, Po, r)
    p = rt
    nu = float(np.sum(r * rt))
    trace = [nu]
    for _ in range(max_iter):
        Ap = bt_matvec(Sd, So, p)
        alpha = nu / float(np.sum(p * Ap))
        r = r - Ap * alpha
        x = x + p * alpha
        rt = bt_matvec(Pd, Po, r)
        nu_prime = float(np.sum(r * rt))
        trace.append(nu_prime)
        if abs(nu_prime) < tol:
            break
        beta = nu_prime / nu
        p = rt + p * beta
        nu = nu_prime
    return x, [abs(t) for t in trace]


def bt_solve_dense(Sd, So, gamma):
    """np.linalg.solve on the assembled S: the reference's method 'S' (:430-436)."""
    from .dense import assemble_bt
    S = assemble_bt(Sd, So)
    return np.linalg.solve(S, gamma.reshape(-1)).reshape(gamma.shape)


def recover(blocks, sch, l, nx):
    """dxu = inv(G) (g - C^T l)  (:449-452) -> dz (N,m) rows [dx_k; du_k] (terminal du = 0)."""
    g = blocks["g"]
    AB, Ghat = sch["AB"], sch["Ghat"]
    rhs = g.copy()
    rhs[:, :nx] -= l
    rhs[:-1] += np.matmul(np.swapaxes(AB, -1, -2), l[1:, :, None])[..., 0]
    return np.matmul(Ghat, rhs[..., None])[..., 0]
