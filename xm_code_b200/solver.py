"""The reference's entry points (XM/src/XM_main.cu:180-310 ``solve``, :312-401 ``solve_rank3``, :35-178 ``solve_rebuttle``) as
a thin Python binding of ONE C-ABI call, ``xm_solve`` (xm_code_b200/csrc/xm_staircase.cu) — the same staircase the compiled
module ``XM`` (XM/src/XM_main.cpp) runs.  Because it lives behind the C-ABI it works on whatever the handle holds: a dense or
block-CSR operator, one GPU or a communicator (then every rank makes the same call and holds the same result; rank 0 writes).

    h = capi.Handle(); h.set_q_dense(Q)                      # or dist.attach(h, N, max_rank) first, on every rank
    out = solver.solve_arrays(h, max_rank=5, tol=1e-6, lam=0.0)
    out["R"], out["s"], out["rank"], out["status"]           # status: 1 certified, 2 max rank reached, -2 line search failed

``solve`` / ``solve_rank3`` / ``solve_rebuttle`` keep the reference's path-based signatures (read ``<path>/Q.bin``, write
``R.bin`` / ``s.bin``).  There is no CPU fallback: the handle is a GPU handle.
"""
from __future__ import annotations

import numpy as np

from . import binio


def solve_arrays(handle, max_rank: int, tol: float, lam: float, max_time: float = 1000.0, certificate_method: str = "auto",
                 s_init=None, rank3_only: bool = False, mode: str | None = None) -> dict:
    """The staircase on a handle whose Q is already set.  certificate_method: "dense" (cuSOLVER syevd on the assembled dual slack:
    one GPU, dense Q), "iterative" (block Davidson on the Q.Y operator: any operator / communicator) or "auto"."""
    mode = mode or ("rank3" if rank3_only else ("rebuttle" if s_init is not None else "full"))
    return handle.solve(max_rank, tol, lam, max_time=max_time, mode=mode, s_init=s_init, cert_method=certificate_method)


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
    out = solve_arrays(h, max_rank, tol, lam, max_time, certificate_method=certificate_method, s_init=s_init, mode=mode)
    if h.comm_info()["rank"] == 0:
        binio.save_matrix_to_bin(dataset_path + "/R.bin", out["R"])
        binio.save_matrix_to_bin(dataset_path + "/s.bin", out["s"][:, None])
    if handle is None:
        h.close()
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
