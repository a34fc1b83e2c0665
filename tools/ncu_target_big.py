"""ncu target: ONE time-capped solve of the bench workload (bench.py's generator and step), so that
`ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:xm_solve_kernel -c 1` measures the DRAM
traffic of the persistent solve kernel of THIS build.  Prints `PRODUCTS <n>` (Q.Y products executed by the captured launch).
  python tools/ncu_target_big.py [max_time_s=0.25]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402
from xm_code_b200 import capi  # noqa: E402

T = float(sys.argv[1]) if len(sys.argv) > 1 else 0.25
N = bench.N_CAMERAS; n3 = 3 * N
Qt, _ = bench.make_problem_device("cuda")
h = capi.Handle(device=0)
h.set_stream(torch.cuda.current_stream().cuda_stream)
h.set_q_dense_dev(n3, Qt.data_ptr(), n3)
torch.cuda.synchronize()
del Qt
torch.cuda.empty_cache()
R0 = bench.identity_start(torch, N, bench.RANK, "cuda"); s0 = torch.ones(N, dtype=torch.float64, device="cuda")
R = torch.empty_like(R0); s = torch.empty_like(s0)
primal, _, st = h.trust_region_dev(bench.RANK, R0.data_ptr(), s0.data_ptr(), R.data_ptr(), s.data_ptr(), lam=bench.LAM, gradtol=bench.GRADTOL, max_time=T)
print(f"PRODUCTS {st['qy_products']} CAMERAS {N} SOLVE_MS {st['solve_ms']:.3f} TCG {st['tcg_iters']}", flush=True)
