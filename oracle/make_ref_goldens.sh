#!/bin/bash
# Generates tests/golden/ref_* from the REFERENCE ITSELF: runs oracle/_ref/xm_ref_harness (the unmodified
# XM/include/XM/trustregion.h, built by oracle/Makefile) on a B200.  Run on the GPU box through gpurun:
#     gpurun --timeout 900 -- 'bash oracle/make_ref_goldens.sh'
# Outputs land in gpurun_out/ref_goldens/ ; copy them to tests/golden/ and commit (done by hand).
set -e
cd "$(dirname "$0")/.."
OUT=gpurun_out/ref_goldens
mkdir -p "$OUT"
H=oracle/_ref/xm_ref_harness
python - <<'PY'
import os, sys, numpy as np
sys.path.insert(0, ".")
from xm_code_b200 import binio, problems
from oracle import xm_oracle as xo
out = "gpurun_out/ref_goldens"
def case(name, Q, r=3, R0=None, s0=None, v=None):
    d = os.path.join(out, name); os.makedirs(d, exist_ok=True)
    binio.save_matrix_to_bin(d + "/Q.bin", Q)
    if R0 is not None: binio.save_matrix_to_bin(d + "/R_ini.bin", R0)
    if s0 is not None: binio.save_matrix_to_bin(d + "/s_ini.bin", s0)
    if v is not None: binio.save_matrix_to_bin(d + "/v_ini.bin", v)
Q1 = binio.load_matrix_from_bin("tests/golden/simple1_Q.bin"); case("simple1", Q1)
Q2 = np.load("tests/golden/simple2_Q_ref.npz")["Q"]; case("simple2", Q2)
Q3, _ = problems.synthetic_dense_q(100, seed=1); case("syn100", Q3)
# rank escalation replay: generic PSD matrix, rank-3 solution from the oracle, escape direction from its certificate
rng = np.random.default_rng(11); N = 30
A = rng.standard_normal((3 * N, 3 * N + 2)); Q4 = A @ A.T / (3 * N)
res3 = xo.trust_region(Q4, xo.identity_init(N, 3), np.ones(N), 0.0, 1e-7)
c = xo.certificate(Q4, xo.from_blocks(res3.Y * res3.s[:, None, None]), 0.0, res3.primal)
Y0 = np.concatenate([res3.Y, np.zeros((N, 3, 1))], axis=2)
v = (c["v"].reshape(N, 3) / res3.s[:, None]).reshape(-1)
case("esc30_r4", Q4, 4, xo.from_blocks(Y0), res3.s, v)
Qb, _ = problems.synthetic_dense_q(1723, seed=0, obs_per_camera=60, n_landmarks=12 * 1723); case("bal1723", Qb)
PY
$H $OUT/simple1 3 1e-16 0.0 1000 > $OUT/simple1/log.txt
$H $OUT/simple2 3 1e-10 0.0 1000 > $OUT/simple2/log.txt
$H $OUT/syn100 3 1e-6 0.0 1000 > $OUT/syn100/log.txt
$H $OUT/esc30_r4 4 1e-7 0.0 1000 1.0 > $OUT/esc30_r4/log.txt
$H $OUT/bal1723 3 1e-6 0.0 1000 0 3 > $OUT/bal1723/log.txt
rm -f $OUT/*/Q.bin $OUT/bal1723/R_ref.bin   # inputs are regenerated from seeds; keep outputs + logs only
tail -n 3 $OUT/*/log.txt
