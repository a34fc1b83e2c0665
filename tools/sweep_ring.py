"""Tuning sweep on the bench workload: ring geometry (XM_TUNE_KC / XM_TUNE_ST env hooks) vs Q.Y time and solve time."""
import os, sys, json, subprocess
code = r'''
import sys, os, json
sys.path.insert(0, ".")
import numpy as np, bench
from xm_code_b200 import capi
from oracle import xm_oracle as xo
Q, _ = bench.make_problem(); N = Q.shape[0] // 3
h = capi.Handle(); h.set_q_dense(Q)
X = np.random.default_rng(0).standard_normal((3 * N, 3)); h.qy(X)
free = h.bench_qy(3, 50); lock = h.bench_qy(3, -50)
R0 = xo.from_blocks(xo.identity_init(N, 3))
h.trust_region(R0, np.ones(N), 0.0, 1e-6)
res = h.trust_region(R0, np.ones(N), 0.0, 1e-6)
st = res.stats
print(json.dumps(dict(kc=os.environ.get("XM_TUNE_KC"), st=os.environ.get("XM_TUNE_ST"), pf=os.environ.get("XM_TUNE_PF"), qy_free_us=free * 1e3, qy_lock_us=lock * 1e3,
      solve_ms=st["solve_ms"], qy_in_solve_us=st["qy_ms"] * 1e3 / st["qy_products"], sync_us=st["sync_ms"] * 1e3 / st["qy_products"],
      per_product_us=st["solve_ms"] * 1e3 / st["qy_products"], its=st["tcg_iters"])))
'''
for pf in [0, 2, 4, 6, 8, 12, 16, 24]:
    env = dict(os.environ, XM_TUNE_PF=str(pf))
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
    print(out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr[-500:], flush=True)
