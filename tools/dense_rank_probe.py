"""Dense Q.Y at ranks 3..20 on the bench workload (lock-step products inside one launch) — A/B of the two-cameras-per-warp sweep
(padded ranks 8 / 10, 256-thread CTAs; XM_TUNE_DENSE_NT=512 restores one camera per warp).  python tools/dense_rank_probe.py [cameras]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from xm_code_b200 import capi, problems  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 13682
n3 = 3 * N
prob = problems.synthetic_sfm_torch(N, n_landmarks=12 * N, obs_per_camera=60, seed=0, device="cuda")
Q = problems.q_from_observations_torch(prob["N"], prob["M"], prob["cam"], prob["lm"], prob["w"], prob["pt"], device="cuda")
peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
for env in (None, "512"):
    if env:
        os.environ["XM_TUNE_DENSE_NT"] = env
    h = capi.Handle(device=0)
    h.set_stream(torch.cuda.current_stream().cuda_stream)
    h.set_q_dense_dev(n3, Q.data_ptr(), n3)
    for r in (3, 5, 6, 8, 10, 12, 16, 20):
        if env and r not in (8, 10):
            continue
        X = torch.randn(r, n3, dtype=torch.float64, device="cuda"); O = torch.empty_like(X)
        h.qy_dev(r, X.data_ptr(), O.data_ptr())
        err = float(((Q @ X.T).T - O).abs().max() / O.abs().max())
        lock = h.bench_qy(r, -8)
        alg = 72.0 * N * N + 48.0 * N * r
        print(json.dumps({"cameras": N, "r": r, "dense_nt_override": env, "ms_per_product_lockstep": lock, "frac_of_hbm_peak": alg / (lock * 1e-3) / 1e9 / peak, "rel_err": err}), flush=True)
    h.close()
