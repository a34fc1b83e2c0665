"""``create_matrix`` with the reference's signature and outputs (utils/creatematrix.py:52-341): observations of the
bipartite (camera, landmark) graph -> ``<output_path>/Q.bin`` (the 3N x 3N SDP data matrix) and ``Abar.bin`` (the map
back to translations and landmarks used by ``recover_XM``).

Same mathematics, different route: the reference forms the least-squares normal equations of the
(N+M) x (N+M-1) bipartite Laplacian densely in M (``V3_bar_F.toarray()``, ``A`` of size (N+M) x 3N) with a two-pass block
solve and a Sherman-Morrison correction on the host.  Here the assembly runs ON THE GPU behind the C-ABI
(``xm_create_matrix``, xm_code_b200/csrc/xm_assemble.cu): the landmark block (diagonal) is eliminated first by one kernel
over the co-observation pairs, leaving one (N-1) x (N-1) Cholesky, a triangular solve and a SYRK:
Q = Q1 - Vbar Lbar^{-1} Vbar^T, Abar = -Lbar^{-1} Vbar^T.  Checked against the reference's own output: Q on SIMPLE2, Q and
Abar on the 24-camera fixture (tests/test_gpu_assemble.py).  No CPU fallback (``problems.q_from_observations`` is the host
restatement the tests and the problem generators use)."""
from __future__ import annotations

import numpy as np

from . import binio


def create_matrix(weight, edges, landmarks, output_path, handle=None):
    """weight: (nobs,), edges: (nobs, 2) 1-based (camera, landmark) like the reference, landmarks: (nobs, 3) camera-frame
    points.  Writes Q.bin and Abar.bin into output_path; returns (Q, Abar).  `handle`: a capi.Handle to assemble on (it keeps Q
    as its operator); default: a temporary handle on device 0."""
    from . import capi
    edges = np.asarray(edges)
    N = int(edges[:, 0].max()); M = int(edges[:, 1].max())
    print(f"M: {M}, N: {N}")
    h = handle or capi.Handle(device=0)
    try:
        Q, Abar, _ = h.create_matrix(N, M, edges[:, 0] - 1, edges[:, 1] - 1, weight, landmarks, want_q=True, want_abar=True)
    finally:
        if handle is None:
            h.close()
    binio.save_matrix_to_bin(output_path + "/Abar.bin", Abar)
    binio.save_matrix_to_bin(output_path + "/Q.bin", Q)
    print(f"Matrix saved to {output_path}/Q.bin\n")
    return Q, Abar
