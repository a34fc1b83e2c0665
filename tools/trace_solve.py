"""Prints the in-kernel timeline (CTA 0 / thread 0) of one tCG iteration of the bench workload (profile mode)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from xm_code_b200 import capi
from oracle import xm_oracle as xo
Q, _ = bench.make_problem()
N = Q.shape[0] // 3
h = capi.Handle(profile=True)
h.set_q_dense(Q)
res = h.trust_region(xo.from_blocks(xo.identity_init(N, 3)), np.ones(N), 0.0, 1e-6)
tr = h.debug_trace()
t0 = tr[0][1]
prev = t0
for tag, t in tr:
    print(f"tag {tag:4d}  t={(t - t0) / 1e3:9.3f} us  dt={(t - prev) / 1e3:8.3f}")
    prev = t
print(res.stats)
