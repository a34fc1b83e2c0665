"""Small ncu target on the bench workload: 3 single Q.Y products through xm_qy_dev (ops kernel, MODE_OUT) and one solve."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from xm_code_b200 import capi
from oracle import xm_oracle as xo
Q, _ = bench.make_problem(); N = Q.shape[0] // 3
h = capi.Handle(qy_variant=int(os.environ.get("XM_QY_VARIANT", "0")))
h.set_q_dense(Q)
X = torch.randn(3, 3 * N, dtype=torch.float64, device="cuda"); O = torch.empty_like(X)
for _ in range(3):
    h.qy_dev(3, X.data_ptr(), O.data_ptr())
torch.cuda.synchronize()
if os.environ.get("XM_NCU_SOLVE", "1") == "1":
    res = h.trust_region(xo.from_blocks(xo.identity_init(N, 3)), np.ones(N), 0.0, 1e-6)
    print(res.stats)
