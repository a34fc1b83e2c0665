"""``recover_XM`` with the reference's signature and return values (utils/recoversolution.py:4-86), computed on the GPU
by ``xm_recover`` (include/xm_b200.h): rank-r -> 3 projection, per-camera scale, anchoring, O(3) projection of every
camera block, and translations / landmarks from ``Abar``.  No CPU fallback: a B200 is required."""
from __future__ import annotations

import numpy as np

from . import capi

_handle = None


def recover_XM(Q, R, s, Abar, lam, handle: "capi.Handle | None" = None):
    """Returns (R_real (3 x 3N), s_real (N,), p_est (3 x M), t_est (3 x N)) like the reference.  Q and lam are accepted for
    signature compatibility: the reference uses them only to print a sub-optimality diagnostic."""
    global _handle
    h = handle
    if h is None:
        if _handle is None:
            _handle = capi.Handle(device=0)
        h = _handle
    out = h.recover(np.asarray(R, dtype=np.float64), np.asarray(s, dtype=np.float64).reshape(-1), Abar)
    ev = out["eigvals"]
    if ev is not None and ev.size > 3:
        if abs(ev[3] / ev[2]) < 1e-3:
            print("Optimal rank is 3")
    if out["negative"] > 0:
        print("warning: some det(R) < 0")
    return out["R"], out["s"], out["p"], out["t"]
