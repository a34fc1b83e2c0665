"""Matrix-free optimality certificate: the reference's ``checkeig`` (XM/include/XM/checkeig.h:42-368)
without its 3N x 3N eigendecomposition, on top of the library's own Q.Y operator.

``checkeig`` solves a least-squares problem for 5N+1 multipliers with Eigen's LSCG on the host, assembles the dual slack
S = Z - sum_j y_j A_j densely on the host and calls ``cusolverDnXsyevd`` on all of it (O(N^3), two 3N x 3N host<->device
copies).  Here:

  * the multipliers are solved in closed form per camera — constraint j only touches its own camera's rows, so the
    normal equations are block diagonal (one 6x6 system for camera 0, one 5x5 per other camera) —
    vectorised over cameras in NumPy (the same algebra ``xm_certify`` runs in C++);
  * S differs from Q only in its 3x3 diagonal blocks, so S x = Q x + B x with B block diagonal: the only O(N^2) work is
    the library's Q.Y product (``xm_qy``; dense or block-CSR, one GPU or a communicator — the operator is collective and
    returns the full product on every rank, so every rank runs the same Lanczos iteration and takes the same decision);
  * lambda_min(S) and its eigenvector come from ARPACK's implicitly restarted Lanczos (``scipy.sparse.linalg.eigsh``)
    on the shifted operator S - c I (c ~ lambda_max: the stopping test is then relative to the spectrum's width instead
    of to lambda_min ~ 0) — typically 500-900 operator applications (measured on the reference's two shipped problems).

The decision rule, dual value and gap are the reference's (checkeig.h:322-368).  A Lanczos Ritz value is an upper bound
of lambda_min that has converged; like every iterative certificate (SE-Sync's included) it cannot prove a lower bound.
The dense path (``xm_certify``) stays the default of the compiled ``XM`` module; this one is what makes the staircase run
on a communicator and beyond a few thousand cameras (``xm_code_b200.solver``).
"""
from __future__ import annotations

import numpy as np
from scipy.sparse.linalg import ArpackNoConvergence, LinearOperator, eigsh

# constraint bases (checkeig.h:71-161): camera 0 — the six symmetric unit matrices; other cameras — five traceless ones
_B0 = np.zeros((6, 3, 3))
_B0[0, 0, 0] = 1.0
_B0[1, 0, 1] = _B0[1, 1, 0] = 0.5
_B0[2, 0, 2] = _B0[2, 2, 0] = 0.5
_B0[3, 1, 1] = 1.0
_B0[4, 1, 2] = _B0[4, 2, 1] = 0.5
_B0[5, 2, 2] = 1.0
_B1 = np.zeros((5, 3, 3))
_B1[0, 0, 0], _B1[0, 1, 1] = 0.5, -0.5
_B1[1, 1, 1], _B1[1, 2, 2] = 0.5, -0.5
_B1[2, 0, 1] = _B1[2, 1, 0] = 0.5
_B1[3, 0, 2] = _B1[3, 2, 0] = 0.5
_B1[4, 1, 2] = _B1[4, 2, 1] = 0.5


def _solve_batched(M, g):
    """Least-squares solution of the (possibly rank-deficient) small SPD systems M y = g, batched (pinv = basic LS solution)."""
    return np.einsum("nij,nj->ni", np.linalg.pinv(M, rcond=1e-13, hermitian=True), g)


def multipliers(sR_blocks, W_blocks):
    """Closed-form least-squares multipliers.  sR_blocks, W_blocks: (N, 3, r) with W = Z sR.
    Returns (Lam (N, 3, 3): sum_j y_j A_j restricted to each camera's diagonal block, y0 (6,) of camera 0)."""
    N = sR_blocks.shape[0]
    Lam = np.zeros((N, 3, 3))
    # camera 0
    C0 = np.einsum("qab,bj->qaj", _B0, sR_blocks[0])                      # images B_q sR_0
    M0 = np.einsum("paj,qaj->pq", C0, C0)[None]
    g0 = np.einsum("paj,aj->p", C0, W_blocks[0])[None]
    y0 = _solve_batched(M0, g0)[0]
    Lam[0] = np.einsum("q,qab->ab", y0, _B0)
    if N > 1:
        C = np.einsum("qab,nbj->nqaj", _B1, sR_blocks[1:])
        M = np.einsum("npaj,nqaj->npq", C, C)
        g = np.einsum("npaj,naj->np", C, W_blocks[1:])
        y = _solve_batched(M, g)
        Lam[1:] = np.einsum("nq,qab->nab", y, _B1)
    return Lam, y0


def certify(qy_op, R, s, lam, primal, tol: float = 1e-7, seed: int = 0, maxiter: int | None = None):
    """qy_op(X) -> Q @ X for X of shape (3N, k), k >= 1 (any backend; with a communicator it must be called collectively).
    R: 3N x r, s: (N,).  Returns dict(certified, min_eig, v (3N,), dual, gap, matvecs, lambda_max)."""
    R = np.asarray(R, dtype=np.float64); s = np.asarray(s, dtype=np.float64).reshape(-1)
    n3, r = R.shape
    N = n3 // 3
    sR = R * np.repeat(s, 3)[:, None]
    xii = np.einsum("ij,ij->i", sR[0::3], sR[0::3])                       # |sR[3i, :]|^2  (ConstructZmatrixKernal :31-40)
    dz = np.zeros(n3); dz[0::3] = 2.0 * lam * (xii - 1.0)
    W = qy_op(sR) + dz[:, None] * sR                                       # Z sR
    Lam, y0 = multipliers(sR.reshape(N, 3, r), W.reshape(N, 3, r))
    B = -Lam
    B[:, 0, 0] += dz[0::3]                                                 # S = Q + blockdiag(B)
    dual = float(y0[0] + y0[3] + y0[5]) + float(np.sum((1.0 - xii * xii) * lam))   # :322-333
    count = [0]

    def apply_S(x):
        x = np.asarray(x, dtype=np.float64).reshape(n3, -1)
        count[0] += x.shape[1]
        return qy_op(x) + np.einsum("nab,nbk->nak", B, x.reshape(N, 3, -1)).reshape(n3, -1)

    rng = np.random.default_rng(seed)
    v0 = rng.standard_normal(n3)
    # spectrum width: a loose largest-eigenvalue estimate is enough to set the shift
    opS = LinearOperator((n3, n3), matvec=lambda x: apply_S(x)[:, 0], dtype=np.float64)
    lmax = float(eigsh(opS, k=1, which="LA", tol=1e-2, v0=v0, return_eigenvectors=False)[0])
    c = 1.05 * max(abs(lmax), 1e-300)
    opT = LinearOperator((n3, n3), matvec=lambda x: apply_S(x)[:, 0] - c * np.asarray(x).reshape(-1), dtype=np.float64)
    ncv = min(n3 - 1, 64)
    # tol is relative to |lambda - c| ~ the spectrum's width: 1e-7 resolves lambda_min to ~1e-7 lambda_max, far below the
    # decision threshold (1e-4 / 1e-3); a tighter tol only buys digits of a near-degenerate cluster at 3-5x the products
    try:
        val, vec = eigsh(opT, k=min(2, n3 - 2), which="SA", tol=tol, v0=v0, ncv=ncv, maxiter=maxiter)
    except ArpackNoConvergence as e:                                       # keep what converged, else one looser retry
        if len(e.eigenvalues) > 0:
            val, vec = e.eigenvalues, e.eigenvectors
        else:
            val, vec = eigsh(opT, k=1, which="SA", tol=100 * tol, v0=v0, ncv=ncv, maxiter=maxiter)
    j = int(np.argmin(val))
    min_eig = float(val[j] + c)
    v = vec[:, j].copy()
    gap = float(primal - dual - 3 * N * min(0.0, min_eig))                 # :334-336
    bound = 1e-3 if N > 2000 else 1e-4                                     # :349-358 (later tiers unreachable, quirk Q5)
    certified = bool((gap / primal < 1e-3) or (min_eig > -bound))          # :360
    return dict(certified=certified, min_eig=min_eig, v=v, dual=dual, gap=gap, matvecs=count[0], lambda_max=lmax, Lam=Lam)


def handle_operator(handle):
    """Q.X through the C-ABI (``xm_qy``): slices of at most XM_MAX_RANK columns, narrow slices padded to the kernel's
    minimum of 3 columns."""
    def qy(X):
        X = np.asarray(X, dtype=np.float64)
        if X.ndim == 1:
            X = X[:, None]
        out = np.empty(X.shape)
        for j0 in range(0, X.shape[1], 20):
            blk = X[:, j0:j0 + 20]
            k = blk.shape[1]
            if k < 3:
                blk = np.pad(blk, ((0, 0), (0, 3 - k)))
            out[:, j0:j0 + k] = handle.qy(np.asfortranarray(blk))[:, :k]
        return out
    return qy
