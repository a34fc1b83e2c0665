"""ncu target: 3 single block-CSR Q.Y products (ops kernel, MODE_OUT) at Erdos-Renyi scale (BASELINE config 5)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from xm_code_b200 import capi, problems
N = int(os.environ.get("XM_NCU_CAMERAS", "100000")); r = int(os.environ.get("XM_NCU_RANK", "10"))
rowptr, col, vals = problems.erdos_renyi_bsr(N, avg_degree=100.0, seed=0)
h = capi.Handle()
h.set_q_bsr(rowptr, col, vals, 3)
X = torch.randn(r, 3 * N, dtype=torch.float64, device="cuda"); O = torch.empty_like(X)
for _ in range(3):
    h.qy_dev(r, X.data_ptr(), O.data_ptr())
torch.cuda.synchronize()
print("nnzb", int(rowptr[-1]), "alg bytes", int(rowptr[-1]) * 132 + 4 * (N + 1) + 2 * 8 * 3 * N * r)
