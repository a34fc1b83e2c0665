"""Probe of a BAL-Final-sized dense problem on one B200 (round-2 sizing of the bench workload):
generation time on the device, one solve with the per-outer-iteration trajectory, Q.Y timings, and (optionally) the
unmodified reference harness on the same Q.   python tools/big_probe.py [cameras] [max_time_s] [ref_max_time_s]"""
import ctypes as C
import json
import os
import re
import shutil
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from xm_code_b200 import capi, problems  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 13682
MAXT = float(sys.argv[2]) if len(sys.argv) > 2 else 60.0
REFT = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0
out = {"cameras": N}
torch.cuda.init()
t0 = time.perf_counter()
prob = problems.synthetic_sfm_torch(N, n_landmarks=12 * N, obs_per_camera=60, seed=0, device="cuda")
torch.cuda.synchronize(); t1 = time.perf_counter()
Q = problems.q_from_observations_torch(prob["N"], prob["M"], prob["cam"], prob["lm"], prob["w"], prob["pt"], device="cuda")
torch.cuda.synchronize(); t2 = time.perf_counter()
out["gen_obs_s"] = t1 - t0; out["assemble_s"] = t2 - t1; out["nobs"] = int(prob["cam"].size); out["landmarks"] = int(prob["M"])
out["peak_mem_gb"] = torch.cuda.max_memory_allocated() / 1e9
print(json.dumps(out), flush=True)
n3 = 3 * N
h = capi.Handle(device=0)
h.set_stream(torch.cuda.current_stream().cuda_stream)
t0 = time.perf_counter()
h.set_q_dense_dev(n3, Q.data_ptr(), n3)
torch.cuda.synchronize(); out["set_q_dev_s"] = time.perf_counter() - t0
R0 = torch.zeros(3, n3, dtype=torch.float64, device="cuda")
for a in range(3):
    R0[a, a::3] = 1.0
s0 = torch.ones(N, dtype=torch.float64, device="cuda")
R = torch.empty_like(R0); s = torch.empty_like(s0)
X = torch.randn(3, n3, dtype=torch.float64, device="cuda"); O = torch.empty_like(X)
for r in (3,):
    h.qy_dev(r, X.data_ptr(), O.data_ptr())
    ref = (Q @ X.T).T
    out[f"qy_err_r{r}"] = float((O - ref).abs().max() / ref.abs().max())
    out[f"qy_ms_free_r{r}"] = h.bench_qy(r, 10)
    out[f"qy_ms_lock_r{r}"] = h.bench_qy(r, -10)
print(json.dumps(out), flush=True)
for ks in (3, 5):          # ring geometry: ksplit = warps sharing one camera's rows (smaller boxes, more stages)
    hk = capi.Handle(device=0, ksplit=ks)
    hk.set_stream(torch.cuda.current_stream().cuda_stream)
    hk.set_q_dense_dev(n3, Q.data_ptr(), n3)
    hk.qy_dev(3, X.data_ptr(), O.data_ptr())
    out[f"qy_ms_lock_r3_ksplit{ks}"] = hk.bench_qy(3, -10)
    out[f"qy_err_ksplit{ks}"] = float((O - ref).abs().max() / ref.abs().max())
    hk.close()
print(json.dumps(out), flush=True)
gt = C.c_double(1e-6); pr = C.c_double(); st = capi.XmStats(); log = (capi.XmLogRec * capi.XM_LOG_CAP)()
rc = h.lib.xm_trust_region_dev(h._h, 3, C.c_void_p(R0.data_ptr()), C.c_void_p(s0.data_ptr()), 0.0, C.byref(gt), 0.0, None, MAXT,
                               C.c_void_p(R.data_ptr()), C.c_void_p(s.data_ptr()), C.byref(pr), C.byref(st), C.cast(log, C.c_void_p))
h._check(rc, "solve")
out.update(solve_ms=st.solve_ms, tcg=st.tcg_iters, outer=st.outer_iters, products=st.qy_products, exit=st.exit_code, primal=pr.value,
           gradnorm=st.gradnorm, qy_ms=st.qy_ms, sync_ms=st.sync_ms, grid=st.grid_ctas)
cum = 0; traj = []
for i in range(st.n_log):
    if i > 0:
        cum += log[i].inner_shown
    traj.append((log[i].k, cum, log[i].loss, log[i].gradnorm))
out["trajectory"] = traj[:: max(1, len(traj) // 60)] + traj[-1:]
print(json.dumps(out), flush=True)
if REFT > 0:
    harness = os.path.join(ROOT, "oracle", "_ref", "xm_ref_harness")
    free_shm = shutil.disk_usage("/dev/shm").free
    d = tempfile.mkdtemp(dir="/dev/shm" if free_shm > 8 * n3 * n3 + (1 << 30) else None)
    t0 = time.perf_counter()
    Qh = Q.cpu().numpy()
    with open(os.path.join(d, "Q.bin"), "wb") as f:
        np.array([n3, n3], dtype=np.int32).tofile(f)
        Qh.tofile(f)
    out["write_q_s"] = time.perf_counter() - t0; out["q_dir"] = d
    del Q; torch.cuda.empty_cache()
    t0 = time.perf_counter()
    p = subprocess.run([harness, d, "3", "1e-6", "0", str(REFT), "0", "1"], capture_output=True, text=True, timeout=1200)
    out["ref_wall_s"] = time.perf_counter() - t0
    m = re.search(r"REFJSON (\{.*\})", p.stdout)
    out["ref"] = json.loads(m.group(1)) if m else p.stdout[-400:] + p.stderr[-400:]
    tot = re.findall(r"Total iteration:\s+(\d+)", p.stdout)
    out["ref_total_iterations"] = [int(x) for x in tot]
    shutil.rmtree(d, ignore_errors=True)
    print(json.dumps(out), flush=True)
