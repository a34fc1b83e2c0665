#!/usr/bin/env python
"""The reference's 2_test_creatematrix.py pipeline against this repo's build, on the SIMPLE2 observations committed under
tests/golden/ (already preprocessed like 2_test_creatematrix.py:29-144):
    observations -> create_matrix (Q.bin, Abar.bin) -> XM.solve -> recover_XM -> rotation error against the ground truth.
Run after `python __graft_entry__.py` on a B200:  python examples/2_pipeline_simple2.py [workdir]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.append(os.path.join(ROOT, "XM", "build"))
sys.path.insert(0, ROOT)

import XM  # noqa: E402

from xm_code_b200.binio import load_matrix_from_bin  # noqa: E402
from xm_code_b200.creatematrix import create_matrix  # noqa: E402
from xm_code_b200.recover import recover_XM  # noqa: E402

work = sys.argv[1] if len(sys.argv) > 1 else "/tmp/xm_simple2"
os.makedirs(work, exist_ok=True)
obs = np.load(os.path.join(ROOT, "tests", "golden", "simple2_obs.npz"))
N = int(obs["N"])

create_matrix(obs["weights"], obs["edges"], obs["pts"], work)        # interface of utils/creatematrix.py:52
lam = 0.0
XM.solve(work, 5, 1e-1, lam, 1000)                                    # 2_test_creatematrix.py:149

Abar = load_matrix_from_bin(work + "/Abar.bin"); R = load_matrix_from_bin(work + "/R.bin")
s = load_matrix_from_bin(work + "/s.bin"); Q = load_matrix_from_bin(work + "/Q.bin")
R_real, s_real, p_est, t_est = recover_XM(Q, R, s, Abar, lam)          # interface of utils/recoversolution.py:4

orig = np.load(os.path.join(ROOT, "tests", "golden", "simple2_frames.npz"))["orig_of_new"]
G = obs["gtR"].reshape(3, -1, 3).transpose(1, 0, 2)[orig]
Rb = R_real.reshape(3, N, 3).transpose(1, 0, 2)
err = [np.degrees(np.arccos(np.clip((np.trace(Rb[i].T @ (G[0] @ G[i].T)) - 1) / 2, -1, 1))) for i in range(N)]
print(f"{N} cameras, {p_est.shape[1]} landmarks; rotation error vs ground truth: median {np.median(err):.3f} deg, max {np.max(err):.3f} deg; "
      f"scales {s_real.min():.4f} .. {s_real.max():.4f}")
