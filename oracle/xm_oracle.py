"""CPU oracle for the XM Burer-Monteiro trust-region path (TEST INFRASTRUCTURE ONLY).

This file is a NumPy restatement of the reference algorithm.  It is the checker for the CUDA path:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it.  The shipped product (``xm_code_b200``, ``XM`` module, ``libxm_b200.so``) never
imports, links or executes anything in ``oracle/``.

Parity pinning: the reference ships no golden vectors.  The oracle is pinned against
outputs of the *reference itself* (``oracle/_ref/xm_ref_harness``: the unmodified ``XM/include/XM/trustregion.h``
compiled by ``oracle/Makefile`` and run on a B200 through gpurun); the per-outer-iteration traces and
final R/s it produced are committed under ``tests/golden/ref_*`` with the generating script
(``oracle/make_ref_goldens.sh``).  See tests/test_oracle_vs_reference.py.

Reference files restated here (all paths relative to /root/reference):
  XM/include/XM/trustregion.h:77-724   XMtrustregion (objective, grad, ehess, ehess2rhess, projection,
                                        retraction, rank-escalation line search, tCG, TR update)
  XM/include/Dense/batchedQR.h:42-67   3-column modified Gram-Schmidt, normalise-then-subtract order
  XM/include/Dense/matdiagmul.h:28-90  per-camera scaling (camera = row/3) and per-camera sums (skip camera 0)
  XM/include/XM/trustregion.h:18-48    scale kernels (retraction exp, lambda terms with the i+1 offsets)
  XM/src/XM_main.cu:180-310            rank staircase (solve), :312-401 solve_rank3, :35-178 solve_rebuttle
  XM/include/XM/checkeig.h:42-368      optimality certificate (multipliers, dual slack, min eig, gap)
  utils/recoversolution.py:4-86        recover_XM (rank-r -> 3, per-camera scale, anchoring, O(3) projection, t/p);
                                        pinned by tests/golden/recover_ref.npz (outputs of the reference function itself)

Layout used here: a point is ``Y`` of shape (N, 3, r) (camera i's 3 x r block, rows orthonormal) and
``s`` of shape (N,) with s[0] == 1 pinned (the reference's ``s_ex``; its ``s`` is the view s_ex[1:]).
The reference stores the same numbers as ``R`` (3N x r column-major) and ``R_T`` (r x 3N column-major);
``Y[i, a, j] == R[3*i + a, j] == R_T[j, 3*i + a]``.  Helpers ``to_blocks`` / ``from_blocks`` convert.
"""
from __future__ import annotations

import math
import time
from dataclasses import dataclass, field

import numpy as np

MAX_INNER_ITER = 1000  # trustregion.h:416
MAX_OUTER_ITER = 1000  # trustregion.h:417


# --------------------------------------------------------------------------- layout helpers
def to_blocks(R: np.ndarray) -> np.ndarray:
    """3N x r (reference ``R``) -> (N, 3, r)."""
    n3, r = R.shape
    return np.ascontiguousarray(R.reshape(n3 // 3, 3, r))


def from_blocks(Y: np.ndarray) -> np.ndarray:
    """(N, 3, r) -> 3N x r."""
    N, _, r = Y.shape
    return Y.reshape(3 * N, r).copy()


def sym3(M: np.ndarray) -> np.ndarray:
    """Batched (A + A^T)/2 — Dense/transpose.h:41-58 symBatchedKernel."""
    return 0.5 * (M + np.swapaxes(M, -1, -2))


# --------------------------------------------------------------------------- elementary operators
def qy(Q: np.ndarray, X: np.ndarray, alpha: float = 1.0) -> np.ndarray:
    """alpha * Q @ X for X in block layout — Dense/matmul.h:42-87 (cublasDgemm, M=K=3N, N=r)."""
    N, _, r = X.shape
    return (alpha * (Q @ X.reshape(3 * N, r))).reshape(N, 3, r)


def objective(Q, Y, s, lam):
    """trustregion.h:162-170 objc: <Q sR, sR> + lam * sum_{i>=1} (s_i^2 - 1)^2."""
    sR = Y * s[:, None, None]
    CsR = qy(Q, sR)
    val = float(np.vdot(CsR, sR))
    reg = float(np.sum((s[1:] * s[1:] - 1.0) ** 2))  # ObjectiveLambdaKernal + Dasum
    return val + lam * reg


def egrad(Q, Y, s, lam, sR=None):
    """trustregion.h:186-194 grad.  Returns (D, G, g) with D = 2 Q sR, G = D*s (per camera),
    g[i] = <D_i, Y_i> + 4 lam (s_i^2-1) s_i for i>=1 and g[0] = 0 (camera 0 has no scale DOF)."""
    if sR is None:
        sR = Y * s[:, None, None]
    D = qy(Q, sR, 2.0)
    G = D * s[:, None, None]
    g = np.einsum("iaj,iaj->i", D, Y)  # matdiagmul.h:61-90 (skips camera 0)
    g = g + 4.0 * lam * (s * s - 1.0) * s  # GradLambdaKernal + Daxpy(4 lam)
    g[0] = 0.0
    return D, G, g


def project(Y, s, G, g):
    """trustregion.h:307-317 projection: G_i - sym(Y_i G_i^T) Y_i ; s^2 * g."""
    S = sym3(np.einsum("iaj,ibj->iab", Y, G))
    rgR = G - np.einsum("iab,ibj->iaj", S, Y)
    rgs = g * s * s
    rgs[0] = 0.0
    return rgR, rgs


def ehess(Q, Y, s, lam, D, P, ps):
    """trustregion.h:227-255 ehess.  D is the stored 2 Q sR ("CsR"); ps[0] must be 0."""
    X = P * s[:, None, None] + Y * ps[:, None, None]  # sRu + suR
    E = qy(Q, X, 2.0)  # CsRu  (the one Q.Y per tCG iteration, :237)
    hr = E * s[:, None, None] + D * ps[:, None, None]
    hs = np.einsum("iaj,iaj->i", E, Y) + np.einsum("iaj,iaj->i", D, P)
    hs = hs + 4.0 * lam * (3.0 * s * s - 1.0) * ps  # HessLambdaKernal
    hs[0] = 0.0
    return hr, hs


def ehess2rhess(Y, s, G, g, hr, hs, P, ps):
    """trustregion.h:277-295 ehess2rhess (the code's '+' at :293 wins over the MATLAB comment's '-')."""
    S = sym3(np.einsum("iaj,ibj->iab", Y, G))  # RTgradR_sym
    T = hr - np.einsum("iab,ibj->iaj", S, P)  # rhr = ehessR - Ru * S
    M = sym3(np.einsum("iaj,ibj->iab", Y, T))  # RTrhr_sym
    rhr = T - np.einsum("iab,ibj->iaj", M, Y)
    rhs = hs * s * s + (ps * s) * g
    rhs[0] = 0.0
    return rhr, rhs


def rhess_vec(Q, Y, s, lam, D, G, g, P, ps):
    """One Riemannian Hessian-vector product = ehess followed by ehess2rhess (trustregion.h:561-563)."""
    hr, hs = ehess(Q, Y, s, lam, D, P, ps)
    return ehess2rhess(Y, s, G, g, hr, hs, P, ps)


def mgs_rows(A: np.ndarray) -> np.ndarray:
    """Dense/batchedQR.h:42-67: for each camera, modified Gram-Schmidt over the 3 rows of the 3 x r block
    (columns of the r x 3 block in the reference's R_T layout): normalise row i, then remove it from rows j>i."""
    Qm = A.copy()
    for i in range(3):
        nrm = np.sqrt(np.sum(Qm[:, i, :] * Qm[:, i, :], axis=1))
        Qm[:, i, :] = Qm[:, i, :] / nrm[:, None]
        for j in range(i + 1, 3):
            d = np.sum(Qm[:, i, :] * Qm[:, j, :], axis=1)
            Qm[:, j, :] = Qm[:, j, :] - d[:, None] * Qm[:, i, :]
    return Qm


def retract(Y, s, etaR, etas, lr=1.0):
    """trustregion.h:341-351 retraction: MGS(Y + lr*eta) ; s * exp(lr * eta_s / s) (pinned s[0] stays 1)."""
    Yn = mgs_rows(Y + lr * etaR)
    sn = s.copy()
    sn[1:] = s[1:] * np.exp(lr * etas[1:] / s[1:])  # positiveManifoldRetractionKernal
    return Yn, sn


def inner(aR, bR, a_s, b_s):
    """trustregion.h:67-74 ProductManifoldInner: two Ddot's (callers pre-divide the scale part by s or s^2)."""
    return float(np.vdot(aR, bR)) + float(np.dot(a_s[1:], b_s[1:]))


# --------------------------------------------------------------------------- trust region
@dataclass
class TRResult:
    Y: np.ndarray
    s: np.ndarray
    primal: float
    gradtol: float  # possibly divided by 10 (quirk Q1, trustregion.h:532-535)
    outer_iters: int = 0
    tcg_iters: int = 0  # "Total iteration" as the reference counts it (sum of i+1)
    qy_products: int = 0
    gradnorm: float = float("nan")
    status: int = 0  # 0 ok, -1 line search failed (primal = -1)
    log: list = field(default_factory=list)  # (k, inner_shown, loss, gradnorm, trstatus, endreason)
    wall_s: float = 0.0


def trust_region(Q, Y0, s0, lam, gradtol, ls_step=0.0, v=None, max_time=1000.0,
                 replicate_stale_sr=True, verbose=False, e_recurrence=False) -> TRResult:
    """Restatement of XMtrustregion (trustregion.h:77-724).

    Q: (3N,3N) dense float64.  Y0: (N,3,r).  s0: (N,) with s0[0]==1.  v: (3N,) escape direction when ls_step!=0.
    ``replicate_stale_sr``: quirk Q3 — after an accepted rank-escalation line search the reference keeps the
    stale ``sR`` (old R) for loss[0], the first gradient and the first CsR (trustregion.h:394-398 vs :422,467,553).
    ``e_recurrence`` (NOT reference behaviour; design study for a tCG iteration with two barriers instead of three,
    DESIGN.md §8): the Euclidean Hessian product E = 2 Q X(p) is not recomputed from the new direction but updated as
    E <- beta E - 2 Q X(r_new), X(r) = s r_R + r_s Y — the product's operand then needs only the new residual, not beta.
    """
    t_start = time.perf_counter()
    Q = np.asarray(Q, dtype=np.float64)
    Y = np.array(Y0, dtype=np.float64, copy=True)
    s = np.array(s0, dtype=np.float64, copy=True)
    s[0] = 1.0
    N, _, o = Y.shape
    dim = N * (3 * o - 6) + N - 1  # :104
    delta_bar = math.sqrt(dim)
    delta = delta_bar / 8.0
    res = TRResult(Y=Y, s=s, primal=0.0, gradtol=gradtol)
    nqy = 0

    sR = Y * s[:, None, None]  # :152

    # ---- rank-escalation line search (:360-408)
    if ls_step != 0:
        f0 = objective(Q, Y, s, lam); nqy += 1
        alpha = float(ls_step)
        dirR = np.zeros_like(Y)
        dirR[:, :, o - 1] = np.asarray(v, dtype=np.float64).reshape(N, 3)  # last column = v (:366)
        Yn = mgs_rows(Y - alpha * dirR)
        f = objective(Q, Yn, s, lam); nqy += 1
        while f > f0:
            alpha = alpha / 2
            Yn = mgs_rows(Y - alpha * dirR)
            f = objective(Q, Yn, s, lam); nqy += 1
            if alpha < 1e-20:
                res.primal = -1.0; res.status = -1; res.qy_products = nqy
                return res
        if f0 - f > 0:
            Y = Yn  # R_T, R <- new ; sR is NOT refreshed in the reference (Q3)
            if not replicate_stale_sr:
                sR = Y * s[:, None, None]
        else:
            res.primal = -1.0; res.status = -1; res.qy_products = nqy
            return res

    loss = np.zeros(MAX_OUTER_ITER + 2)
    gradnorm = np.zeros(MAX_OUTER_ITER + 2)
    # objc(sR, s) with the (possibly stale) sR  (:422)
    loss[0] = float(np.vdot(qy(Q, sR), sR)) + lam * float(np.sum((s[1:] ** 2 - 1.0) ** 2)); nqy += 1

    endreason = 6
    trstatus = 4
    shrink_count = 0
    totalite = 0
    i = 0
    k = 0
    for k in range(MAX_OUTER_ITER):
        vR = np.zeros_like(Y); vs = np.zeros(N)
        hvR = np.zeros_like(Y); hvs = np.zeros(N)
        bestY = Y.copy(); bests = s.copy(); bestloss = loss[k]
        # grad(R, s_ex, sR) — uses the stored sR (stale after an accepted line search), but the current R (:467)
        D = qy(Q, sR, 2.0); nqy += 1
        G = D * s[:, None, None]
        g = np.einsum("iaj,iaj->i", D, Y) + 4.0 * lam * (s * s - 1.0) * s
        g[0] = 0.0
        rgR, rgs = project(Y, s, G, g)
        rR = rgR.copy(); rs = rgs.copy()
        pR = -rgR; ps = -rgs
        rsds = np.zeros(N); rsds[1:] = rs[1:] / s[1:]
        rdotr = inner(rR, rR, rsds, rsds)  # :484
        gradnorm[k] = math.sqrt(rdotr)
        res.log.append((k, i + 1, float(loss[k]), float(gradnorm[k]), trstatus if k > 0 else 0,
                        endreason if k > 0 else 0))
        if verbose:
            print(f"{k}   {i + 1}   {loss[k]:1.3e}   {gradnorm[k]:1.3e}  ts={trstatus} er={endreason}")
        if endreason == 5:  # :527
            break
        if gradnorm[k] < gradtol:  # :532
            gradtol /= 10
            break
        if int(time.perf_counter() - t_start) > max_time:  # :538-543 (integer seconds)
            break
        endreason = 6
        trstatus = 4
        vdotv = 0.0; vdotp = 0.0; pdotp = rdotr
        # CsR = 2 Q sR (:553) — same product as D above (same sR); the reference recomputes it.
        nqy += 1
        i = 0
        E_rec = None
        for i in range(MAX_INNER_ITER):
            if e_recurrence:
                if E_rec is None:
                    E_rec = qy(Q, pR * s[:, None, None] + Y * ps[:, None, None], 2.0); nqy += 1
                hr = E_rec * s[:, None, None] + D * ps[:, None, None]
                hs = np.einsum("iaj,iaj->i", E_rec, Y) + np.einsum("iaj,iaj->i", D, pR) + 4.0 * lam * (3.0 * s * s - 1.0) * ps
                hs[0] = 0.0
            else:
                hr, hs = ehess(Q, Y, s, lam, D, pR, ps); nqy += 1
            rhr, rhs = ehess2rhess(Y, s, G, g, hr, hs, pR, ps)
            rhsds = np.zeros(N); rhsds[1:] = rhs[1:] / (s[1:] ** 2)
            alpha = rdotr / inner(pR, rhr, ps, rhsds)  # :566
            if rdotr < 1e-15:  # :572
                endreason = 5
                break
            if alpha <= 0 or (vdotv + 2 * alpha * vdotp + alpha * alpha * pdotp > delta * delta):
                tau = (-vdotp + math.sqrt(vdotp * vdotp + pdotp * (delta * delta - vdotv))) / pdotp
                vR += tau * pR; vs += tau * ps
                hvR += tau * rhr; hvs += tau * rhs
                endreason = 1 if alpha <= 0 else 2
                break
            vR += alpha * pR; vs += alpha * ps
            rR += alpha * rhr; rs += alpha * rhs
            hvR += alpha * rhr; hvs += alpha * rhs
            rsds[1:] = rs[1:] / s[1:]
            rdotr_new = inner(rR, rR, rsds, rsds)  # :626
            if math.sqrt(rdotr_new) < gradnorm[k] * min(gradnorm[k], 0.1):  # :627
                endreason = 3
                break
            beta = rdotr_new / rdotr
            if e_recurrence:      # the product of this iteration: operand from the NEW RESIDUAL only
                E_rec = beta * E_rec - qy(Q, rR * s[:, None, None] + Y * rs[:, None, None], 2.0); nqy += 1
            pR = beta * pR - rR
            ps = beta * ps - rs
            vdotv, vdotp, pdotp = (vdotv + 2 * alpha * vdotp + alpha * alpha * pdotp,
                                   beta * (vdotp + alpha * pdotp),
                                   beta * beta * pdotp + rdotr_new)  # :642-644
            rdotr = rdotr_new
        else:
            i = MAX_INNER_ITER  # loop ran to completion: the reference's i == max_inner_iter
        totalite += i + 1
        vsds = np.zeros(N); vsds[1:] = vs[1:] / (s[1:] ** 2)
        loss_qu = inner(vR, hvR, vsds, hvs) / 2 + inner(vR, rgR, vsds, rgs)  # :668
        if loss_qu >= 0:  # :669
            break
        Yn, sn = retract(Y, s, vR, vs, 1.0)  # :673
        Y = Yn; s = sn
        sR = Y * s[:, None, None]  # :677
        loss[k + 1] = float(np.vdot(qy(Q, sR), sR)) + lam * float(np.sum((s[1:] ** 2 - 1.0) ** 2)); nqy += 1
        rou = (loss[k + 1] - loss[k]) / loss_qu
        if rou < 0.25:
            delta = delta * 0.25; trstatus = 1; shrink_count += 1
        elif rou > 0.75 and endreason <= 2:
            delta = min(delta * 2, delta_bar); trstatus = 2; shrink_count = 0
        else:
            shrink_count = 0
        if shrink_count > 3:
            delta = delta * 1e-3; shrink_count = 0
            if delta < 1e-20:
                break
        if loss[k + 1] > bestloss or rou < 0.1:  # :702 reject
            Y = bestY; s = bests
            loss[k + 1] = bestloss
            sR = Y * s[:, None, None]
            trstatus = 3
    else:
        k = MAX_OUTER_ITER  # Q2: the reference reads loss[1000] out of bounds here
    res.Y = Y; res.s = s
    res.primal = float(loss[k])  # :715
    res.gradtol = gradtol
    res.outer_iters = k
    res.tcg_iters = totalite
    res.qy_products = nqy
    res.gradnorm = float(gradnorm[k]) if k < len(gradnorm) else float("nan")
    res.wall_s = time.perf_counter() - t_start
    return res


# --------------------------------------------------------------------------- certificate (checkeig.h:42-368)
def _constraint_columns(sR: np.ndarray):
    """Columns A_j @ sR of the multiplier operator (checkeig.h:71-161), dense (3N*r) x (5N+1), vec = column-major."""
    n3, o = sR.shape
    N = n3 // 3
    cols = []

    def col(entries):  # entries: list of (row, source_row, coeff)
        c = np.zeros((n3, o))
        for row, src, w in entries:
            c[row, :] += w * sR[src, :]
        return c.reshape(-1, order="F")

    for a in range(3):
        for b in range(a, 3):
            if a == b:
                cols.append(col([(a, a, 1.0)]))
            else:
                cols.append(col([(a, b, 0.5), (b, a, 0.5)]))
    for i in range(1, N):
        a, b, c = 3 * i, 3 * i + 1, 3 * i + 2
        cols.append(col([(a, a, 0.5), (b, b, -0.5)]))
        cols.append(col([(b, b, 0.5), (c, c, -0.5)]))
        cols.append(col([(a, b, 0.5), (b, a, 0.5)]))
        cols.append(col([(a, c, 0.5), (c, a, 0.5)]))
        cols.append(col([(b, c, 0.5), (c, b, 0.5)]))
    return np.stack(cols, axis=1)


def certificate(Q, sR, lam, primal):
    """checkeig (checkeig.h:42-368) with a dense least-squares in place of Eigen's LSCG (same minimiser).

    Returns dict(certified, min_eig, v, dual, gap, y)."""
    n3, o = sR.shape
    N = n3 // 3
    Z = np.array(Q, dtype=np.float64, copy=True)
    for i in range(N):  # ConstructZmatrixKernal :31-40
        Z[3 * i, 3 * i] += 2.0 * lam * (float(np.dot(sR[3 * i], sR[3 * i])) - 1.0)
    right = (Z @ sR).reshape(-1, order="F")
    A = _constraint_columns(sR)
    y, *_ = np.linalg.lstsq(A, right, rcond=None)
    S = Z.copy()
    cnt = 0
    for a in range(3):
        for b in range(a, 3):
            if a == b:
                S[a, a] -= y[cnt]
            else:
                S[a, b] -= 0.5 * y[cnt]; S[b, a] -= 0.5 * y[cnt]
            cnt += 1
    for i in range(1, N):
        a, b, c = 3 * i, 3 * i + 1, 3 * i + 2
        S[a, a] -= 0.5 * y[cnt]; S[b, b] += 0.5 * y[cnt]; cnt += 1
        S[b, b] -= 0.5 * y[cnt]; S[c, c] += 0.5 * y[cnt]; cnt += 1
        S[a, b] -= 0.5 * y[cnt]; S[b, a] -= 0.5 * y[cnt]; cnt += 1
        S[a, c] -= 0.5 * y[cnt]; S[c, a] -= 0.5 * y[cnt]; cnt += 1
        S[b, c] -= 0.5 * y[cnt]; S[c, b] -= 0.5 * y[cnt]; cnt += 1
    W, V = np.linalg.eigh(0.5 * (S + S.T))
    min_eig = float(W[0])
    dual = float(y[0] + y[3] + y[5])
    xii = np.array([float(np.dot(sR[3 * i], sR[3 * i])) for i in range(N)])
    dual += float(np.sum((1.0 - xii * xii) * lam))
    gap = primal - dual - 3 * N * min(0.0, min_eig)
    bound = 1e-3 if N > 2000 else 1e-4  # :349-358 (later tiers unreachable, Q5)
    certified = (gap / primal < 1e-3) or (min_eig > -bound)
    return dict(certified=bool(certified), min_eig=min_eig, v=V[:, 0].copy(), dual=dual, gap=float(gap), y=y)


# --------------------------------------------------------------------------- staircase (XM_main.cu:180-310)
def identity_init(N: int, r: int = 3) -> np.ndarray:
    """XM_main.cu:230-236: every camera starts at [I_3 0]."""
    Y = np.zeros((N, 3, r))
    for a in range(3):
        Y[:, a, a] = 1.0
    return Y


def solve(Q, max_rank, tol, lam, max_time=1000.0, rank3_only=False, Y_init=None, s_init=None, verbose=False):
    """XM.solve / solve_rank3 / solve_rebuttle staircase.  Returns dict(R (3N x r), s (N,), rank, status, trace)."""
    Q = np.asarray(Q, dtype=np.float64)
    N = Q.shape[0] // 3
    o = 3
    gradtol = tol
    s0 = np.ones(N) if s_init is None else np.array(s_init, dtype=np.float64)
    Y0 = identity_init(N, 3)
    v = np.zeros(3 * N)
    status = 0
    trace = []
    while o <= max_rank:
        if o == 3:
            # solve_rebuttle loads R_ini but then overwrites R0 with the identity at o==3 (XM_main.cu:95-103)
            res = trust_region(Q, identity_init(N, 3), s0, lam, gradtol, 0.0, v, max_time, verbose=verbose)
        else:
            res = trust_region(Q, Y0, s0, lam, gradtol, 1.0, v, max_time, verbose=verbose)
        gradtol = res.gradtol
        trace.append(res)
        if res.primal < 0:
            status = -2
            o += 1
            break
        if rank3_only:
            Y0, s0 = res.Y, res.s
            o += 1
            break
        sR = from_blocks(res.Y * res.s[:, None, None])
        cert = certificate(Q, sR, lam, res.primal)
        if cert["certified"]:
            o += 1
            Y0, s0 = res.Y, res.s
            status = 1
            break
        elif o < max_rank:
            Y0 = np.concatenate([res.Y, np.zeros((N, 3, 1))], axis=2)  # zero-padded new column (:265-269)
            s0 = res.s
            v = cert["v"].copy()
            v = (v.reshape(N, 3) / res.s[:, None]).reshape(-1)  # DecentDirectionKernal (XM_main.cu:8-16)
        else:
            Y0, s0 = res.Y, res.s
            status = 2
        o += 1
    return dict(R=from_blocks(Y0), s=np.asarray(s0), rank=o - 1, status=status, trace=trace)


# --------------------------------------------------------------------------- solution recovery (utils/recoversolution.py:4-86)
def recover(R, s, Abar=None):
    """recover_XM restated.  R: 3N x r, s: (N,), Abar: (N+M-1) x 3N or None.
    Returns dict(R (3 x 3N), s (N,), t (3 x N), p (3 x M), eigvals (descending, of (sR)(sR)^T; None when r == 3),
    negative = number of cameras whose projected block had det < 0 before the global sign decision)."""
    R = np.asarray(R, dtype=np.float64); s = np.asarray(s, dtype=np.float64).reshape(-1)
    N = s.shape[0]
    sR = R * np.repeat(s, 3)[:, None]                                   # :7-9
    eigvals = None
    if R.shape[1] > 3:                                                   # :11-23
        w, V = np.linalg.eigh(sR @ sR.T)
        idx = np.argsort(w)[::-1]
        w, V = w[idx], V[:, idx]
        sR_real = (V[:, :3] * np.sqrt(w[:3])).T
        eigvals = w
    else:                                                                # :32-37
        sR_real = sR.T.copy()
    blocks = sR_real.reshape(3, N, 3).transpose(1, 0, 2)                 # blocks[i] = sR_real[:, 3i:3i+3]
    s_real = np.linalg.norm(blocks, axis=(1, 2)) / np.sqrt(3.0)          # :42-44
    Rb = blocks / s_real[:, None, None]
    Rb = Rb[0].T @ Rb                                                    # anchoring :47-48
    U, _, Vt = np.linalg.svd(Rb)
    UV = U @ Vt
    negative = int(np.sum(np.linalg.det(UV) < 0))                        # :50-55
    if negative > N / 2:                                                 # :62-63
        UV = -UV                                                         # polar(-M) = -polar(M)
    Rb = UV                                                              # :65-73 (both branches assign U Vt)
    R_real = Rb.transpose(1, 0, 2).reshape(3, 3 * N)
    out = dict(R=R_real, s=s_real, eigvals=eigvals, negative=negative, t=None, p=None)
    if Abar is not None:
        sR_out = (Rb * s_real[:, None, None]).transpose(1, 0, 2).reshape(3, 3 * N)
        y = np.hstack([np.zeros((3, 1)), (np.asarray(Abar) @ sR_out.T).T])   # :76-82
        out["t"] = y[:, :N]; out["p"] = y[:, N:]
    return out


def observation_errors(edges, landmarks, weights, R_real, s_real, t_est, p_est):
    """Per-observation weighted squared residual of the XM^2 outlier cut, as written in the reference's pipeline script
    (3_test_colmap_glomap.py:304-316).  edges: (n, 2) 1-based (camera, landmark)."""
    N = np.asarray(s_real).shape[0]
    src = np.asarray(edges)[:, 0] - 1; dst = np.asarray(edges)[:, 1] - 1
    Rm = np.asarray(R_real).reshape(3, N, 3).transpose(1, 0, 2)[src]
    moved = np.asarray(s_real)[src, None] * np.einsum("nij,nj->ni", Rm, landmarks) + np.asarray(t_est)[:, src].T
    diff = np.asarray(p_est)[:, dst].T - moved
    return np.asarray(weights) * np.sum(diff ** 2, axis=1)


# --------------------------------------------------------------------------- .bin wire format (utils/io.py:17-58)
def load_bin(path):
    with open(path, "rb") as f:
        rows = int.from_bytes(f.read(4), "little"); cols = int.from_bytes(f.read(4), "little")
        data = np.fromfile(f, dtype=np.float64, count=rows * cols)
    return data.reshape((rows, cols), order="F")


def save_bin(path, M):
    M = np.asarray(M, dtype=np.float64)
    if M.ndim == 1:
        M = M[:, None]
    with open(path, "wb") as f:
        f.write(int(M.shape[0]).to_bytes(4, "little")); f.write(int(M.shape[1]).to_bytes(4, "little"))
        M.T.tofile(f)  # C-order dump of M^T == column-major M
