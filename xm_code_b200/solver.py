"""The rank staircase of the reference's entry points (XM/src/XM_main.cu:180-310 ``solve``, :312-401 ``solve_rank3``,
:35-178 ``solve_rebuttle``) in Python over the C-ABI — the same control flow as the compiled module ``XM``
(XM/src/XM_main.cpp), plus what that one cannot do: run on a communicator (one solve partitioned over several GPUs) and
certify without a dense eigendecomposition (``certificate.certify``: Lanczos on the library's own Q.Y operator).

    h = capi.Handle(); h.set_q_dense(Q)                      # or dist.attach(h, N, max_rank) first, on every rank
    out = solver.solve_arrays(h, max_rank=5, tol=1e-6, lam=0.0)
    out["R"], out["s"], out["rank"], out["status"]           # status: 1 certified, 2 max rank reached, -2 line search failed

``solve`` / ``solve_rank3`` / ``solve_rebuttle`` keep the reference's path-based signatures (read ``<path>/Q.bin``, write
``R.bin`` / ``s.bin``).  With a communicator every rank makes the same calls and holds the same result; rank 0 writes.
There is no CPU fallback: the handle is a GPU handle.
"""
from __future__ import annotations

import numpy as np

from . import binio, certificate


def _identity_init(N: int) -> np.ndarray:
    R0 = np.zeros((3 * N, 3), order="F")                      # XM_main.cu:230-236: every camera starts at [I_3]
    for a in range(3):
        R0[a::3, a] = 1.0
    return R0


def _certify(handle, R, s, lam, primal, method):
    if method == "dense":
        c = handle.certify(R, s, lam, primal)
        return c["certified"], c["min_eig"], c["v"], c
    c = certificate.certify(certificate.handle_operator(handle), R, s, lam, primal)
    return c["certified"], c["min_eig"], c["v"], c


def solve_arrays(handle, max_rank: int, tol: float, lam: float, max_time: float = 1000.0, certificate_method: str = "auto",
                 s_init=None, rank3_only: bool = False, verbose: bool = False) -> dict:
    """The staircase on a handle whose Q is already set.  certificate_method: "dense" (xm_certify: cuSOLVER syevd, single GPU,
    dense Q), "lanczos" (matrix-free, any operator / communicator) or "auto" (dense for <= 2000 cameras on one GPU)."""
    N = handle.N
    info = handle.comm_info() if hasattr(handle, "comm_info") else {"world": 1}
    method = certificate_method
    if method == "auto":
        method = "dense" if (info.get("world", 1) == 1 and N <= 2000 and not getattr(handle, "is_bsr", False)) else "lanczos"
    o = 3
    gradtol = float(tol)
    s0 = np.ones(N) if s_init is None else np.array(s_init, dtype=np.float64).reshape(-1)
    R0 = _identity_init(N)
    v = np.zeros(3 * N)
    status = 0
    trace = []
    cert = None
    while o <= max_rank or rank3_only:
        if verbose:
            print("+++++++++++++++++++++++++++++++++\nSolve TR with Rank   %d\n+++++++++++++++++++++++++++++++++" % o)
        if o == 3:
            # solve_rebuttle loads R_ini but the reference then overwrites R0 with the identity at o == 3 (XM_main.cu:95-103)
            res = handle.trust_region(_identity_init(N), s0, lam=lam, gradtol=gradtol, ls_step=0.0, v=None, max_time=max_time)
        else:
            res = handle.trust_region(R0, s0, lam=lam, gradtol=gradtol, ls_step=1.0, v=v, max_time=max_time)
        gradtol = res.gradtol                                      # quirk Q1: tightened by every small-gradient exit
        trace.append(res)
        if rank3_only:
            R0, s0 = res.R, res.s
            o += 1
            break
        if res.primal < 0:                                         # line search failed (XM_main.cu:244-247)
            status = -2
            o += 1
            break
        certified, min_eig, vec, cert = _certify(handle, res.R, res.s, lam, res.primal, method)
        if verbose:
            print("The min eig is: %1.3e" % min_eig)
        if certified:
            o += 1
            R0, s0 = res.R, res.s
            status = 1
            break
        elif o < max_rank:
            R0 = np.zeros((3 * N, o + 1), order="F")               # zero-padded new column (:265-269)
            R0[:, :o] = res.R
            s0 = res.s
            v = (np.asarray(vec).reshape(N, 3) / res.s[:, None]).reshape(-1)     # DecentDirectionKernal (XM_main.cu:8-16)
        else:
            R0, s0 = res.R, res.s
            status = 2
        o += 1
    return dict(R=np.asarray(R0), s=np.asarray(s0), rank=o - 1, status=status, trace=trace, certificate=cert,
                certificate_method=method)


def _run(dataset_path, max_rank, tol, lam, max_time, mode, handle=None, certificate_method="auto"):
    from . import capi
    Q = binio.load_matrix_from_bin(dataset_path + "/Q.bin")
    if Q.shape[0] != Q.shape[1] or Q.shape[0] % 3:
        raise ValueError("Q.bin missing or not a square 3N x 3N matrix")
    h = handle or capi.Handle(device=0)
    h.set_q_dense(Q)
    s_init = None
    if mode == "rebuttle":
        try:
            s_init = binio.load_matrix_from_bin(dataset_path + "/s_ini.bin")[:, 0]
        except OSError:
            s_init = None
    out = solve_arrays(h, 3 if mode == "rank3" else max_rank, tol, lam, max_time, certificate_method=certificate_method,
                       s_init=s_init, rank3_only=(mode == "rank3"))
    rank = h.comm_info()["rank"] if hasattr(h, "comm_info") else 0
    if rank == 0:
        binio.save_matrix_to_bin(dataset_path + "/R.bin", out["R"])
        binio.save_matrix_to_bin(dataset_path + "/s.bin", out["s"][:, None])
    return out


def solve(dataset_path, max_rank, tol, lam, max_time, handle=None, certificate_method="auto"):
    """XM.solve(dataset_path, max_rank, tol, lam, max_time): reads Q.bin, writes R.bin / s.bin."""
    _run(dataset_path, max_rank, tol, lam, max_time, "full", handle, certificate_method)


def solve_rank3(dataset_path, max_rank, tol, lam, max_time, handle=None):
    """XM.solve_rank3: one rank-3 trust-region solve, no certificate."""
    _run(dataset_path, max_rank, tol, lam, max_time, "rank3", handle)


def solve_rebuttle(dataset_path, max_rank, tol, lam, max_time, handle=None, certificate_method="auto") -> int:
    """XM.solve_rebuttle: returns 1 certified, 2 max rank reached uncertified, -2 line-search failure."""
    return _run(dataset_path, max_rank, tol, lam, max_time, "rebuttle", handle, certificate_method)["status"]
