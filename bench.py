#!/usr/bin/env python
"""bench.py — XM Burer-Monteiro trust-region throughput on B200 (driver contract in the task statement).

Workload (BASELINE.json config 4, the bandwidth-bound regime the path is built for): BAL-Final-13682-shaped synthetic dense Q
(13 682 cameras, 3N = 41 046, Q = 13.5 GB FP64), built ON THE DEVICE by a plain-torch generator shared by both arms
(xm_code_b200/problems.py: synthetic_sfm_torch + q_from_observations_torch — deterministic, ~4 s).  One "step" = one
XMtrustregion-equivalent call at rank 3 from the reference's identity start with gradtol 1e-6 and the reference's own time
limit `maxtime` = 4 s (XM/include/XM/trustregion.h:77, time exit :538-543).  A full solve to the KKT tolerance is ~7 000 tCG
iterations here (16 s on one B200, ~55 s for the reference): 25 of those per arm do not fit a bench run, so every step runs
the first 4 s of that solve (the whole solve where it takes less: 2.1 s on 8 GPUs) and the metric is the rate, tCG iterations
per second (each iteration = one Q.Y over all of Q + the fused per-camera work).  The full time-to-KKT is measured once per
run and reported in `result.full_solve`; `result.bsr_er100k` is BASELINE config 5 (block-CSR) on the same GPUs.
XM_BENCH_CAMERAS=<n> changes the camera count; below 4000 cameras (e.g. 1723 = BAL-Ladybug, the round-1 workload, 214 MB)
the NumPy generator is used and a step is a FULL solve.

--gpus N > 1 (torchrun, one process per GPU): the SAME solve partitioned by camera over the N GPUs ("scaling": "strong") —
rank k holds the rows of Q of its cameras; per tCG iteration the persistent kernels exchange the operand rows and the
reduction scalars by peer-mapped stores over NVLink (xm_code_b200/dist.py, include/xm_b200.h); no NCCL call on the path.

  value : device-resident (Q, R0, s0 already in HBM; CUDA events on the launch stream)
  e2e   : through the C-ABI with HOST buffers — xm_set_q_dense[_slab] (pinned H2D of Q + re-layout) + xm_trust_region
          (H2D of R0/s0, solve, D2H of R/s) inside the timed region
  roofline      : the persistent solve kernel; algorithmic bytes 8 rows 3N + 8 (3N + rows) r per product
  cpu_baseline  : the compiled C + OpenMP oracle (port; oracle/xm_oracle_c.c) on all host cores, bounded sample
  latency_regime: BAL-Ladybug-1723 (214 MB, the round-1 workload): full solves, latency-bound        (N = 1 only)
  --impl reference : the UNMODIFIED reference trustregion.h (oracle/_ref/xm_ref_harness, cuBLAS path) on the same Q on
          the same GPU — the reference has no CPU implementation of this path; falls back to the oracle port on the
          host when the harness or a GPU is missing.
"""
from __future__ import annotations

import argparse
import json
import os
import re
import shutil
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_CAMERAS = int(os.environ.get("XM_BENCH_CAMERAS", "13682"))
RANK = 3
GRADTOL = 1e-6
LAM = 0.0
BIG = N_CAMERAS >= 4000
STEP_SECONDS = float(os.environ.get("XM_BENCH_STEP_SECONDS", "4")) if BIG else 1000.0
OBS_PER_CAMERA = 60
NAME = {13682: "BAL-Final-13682", 1778: "BAL-Venice-1778", 1723: "BAL-Ladybug-1723"}.get(N_CAMERAS, f"BAL-shaped-{N_CAMERAS}")
CONFIG = {
    "workload": f"{NAME}-shaped synthetic dense Q (3N={3 * N_CAMERAS}, {72 * N_CAMERAS ** 2 / 1e9:.2f} GB FP64 > L2), rank-3 XMtrustregion call from "
                "identity, gradtol 1e-6" + (f", time limit {STEP_SECONDS:g} s per step (the reference's maxtime exit)" if BIG else ", full solve per step"),
    "cameras": N_CAMERAS, "rank": RANK, "gradtol": GRADTOL, "lam": LAM, "step_max_time_s": STEP_SECONDS if BIG else None,
    "generator": ("problems.synthetic_sfm_torch + q_from_observations_torch (seed 0, 60 observations per camera, 12 N landmarks), on the device"
                  if BIG else "problems.synthetic_dense_q (seed 0, 60 observations per camera, 12 N landmarks), NumPy on the host"),
    "l2": "inputs larger than L2 (126 MB): every Q.Y product streams its rows of Q from HBM",
}


def make_problem_host(n=N_CAMERAS):
    from xm_code_b200 import problems
    Q, prob = problems.synthetic_dense_q(n, seed=0, obs_per_camera=OBS_PER_CAMERA, n_landmarks=12 * n)
    return Q, prob


def make_problem_device(device):
    """Q as a torch tensor on `device` (symmetric: its memory is the column-major matrix too) + the observation lists."""
    import torch
    from xm_code_b200 import problems
    if BIG:
        prob = problems.synthetic_sfm_torch(N_CAMERAS, n_landmarks=12 * N_CAMERAS, obs_per_camera=OBS_PER_CAMERA, seed=0, device=device)
        Q = problems.q_from_observations_torch(prob["N"], prob["M"], prob["cam"], prob["lm"], prob["w"], prob["pt"], device=device)
        return Q, prob
    Qh, prob = make_problem_host()
    return torch.from_numpy(np.ascontiguousarray(Qh)).to(device), prob


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        self.index = index; self.proc = None; self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv"); os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index), "-lms", "100"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if not self.proc:
            return None
        try:
            self.proc.terminate(); self.proc.wait(timeout=5)
        except Exception:
            pass
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 7:
                    continue
                try:
                    sm.append(float(f[0])); mx.append(float(f[1]))
                except ValueError:
                    continue
                for n, v in zip(names, f[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            os.unlink(self.path)
        except Exception:
            return None
        if not sm:
            return None
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak_gbs():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


SOLVE_KERNEL_SOURCES = ("xm_device.cuh", "xm_solve.cuh", "xm_inst.cu", "xm_capi.cu", "xm_host.h", "../../include/xm_b200.h")


def source_hash():
    """Hash of the sources that define the persistent solve kernel and its launch plan: identifies the build an ncu capture belongs to."""
    import hashlib
    hsh = hashlib.sha256()
    d = os.path.join(ROOT, "xm_code_b200", "csrc")
    for f in SOLVE_KERNEL_SOURCES:
        hsh.update(open(os.path.join(d, f), "rb").read())
    return hsh.hexdigest()[:16]


def measured_traffic(n_cameras):
    """dram bytes per Q.Y product of the persistent solve kernel from an ncu capture of THIS build (tools/ncu_traffic.py writes
    profiles/r02_solve_traffic.json with the hash of the library's sources); None when there is no capture for this workload / build."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "r02_solve_traffic.json")))
        sha = source_hash()
        for rec in t["captures"]:
            if rec["cameras"] == n_cameras and rec["src_sha16"] == sha:
                return rec["dram_bytes_per_product"], f"ncu dram__bytes_read.sum + dram__bytes_write.sum of one solve launch of this build ({rec['source']}), per product"
    except Exception:
        pass
    return None, "no ncu capture of this build for this workload (tools/ncu_traffic.py)"


def oracle_sample(Q, seconds):
    """cpu_baseline: the CPU oracle on ALL host cores for ~`seconds` of the same solve (bounded sample).  Preferred: the
    compiled C + OpenMP twin (oracle/xm_oracle_c.c, built here for this machine's CPU); fallback: the NumPy oracle.
    Returns (tCG it/s, seconds, result, description, threads)."""
    from oracle import xm_oracle as xo
    N = Q.shape[0] // 3
    try:
        from oracle import xm_oracle_c as xc
        xc.load()
        Qc = np.ascontiguousarray(Q)
        t0 = time.perf_counter()
        res = xc.trust_region(Qc, xo.identity_init(N, 3), np.ones(N), LAM, GRADTOL, max_time=seconds)
        dt = time.perf_counter() - t0
        return res.tcg_iters / dt, dt, res, f"C + OpenMP oracle (oracle/xm_oracle_c.c, gcc -O3 -march=native) on {xc.cpu_model()}", xc.num_threads()
    except Exception as e:  # noqa: BLE001  (no gcc on the box: keep a baseline anyway)
        t0 = time.perf_counter()
        res = xo.trust_region(Q, xo.identity_init(N, 3), np.ones(N), LAM, GRADTOL, max_time=seconds)
        dt = time.perf_counter() - t0
        return res.tcg_iters / dt, dt, res, f"NumPy oracle (OpenBLAS dgemm Q.Y; C oracle unavailable: {type(e).__name__})", os.cpu_count()


def identity_start(torch, N, r, device):
    R0 = torch.zeros(r, 3 * N, dtype=torch.float64, device=device)      # memory == 3N x r column-major
    for a in range(3):
        R0[a, a::3] = 1.0
    return R0


def latency_regime(torch, capi, stream, peak, steps=3):
    """BAL-Ladybug-1723 (the round-1 workload: 214 MB of Q, latency-bound): full solves, device-resident."""
    n = 1723
    Qh, _ = make_problem_host(n)
    n3 = 3 * n
    h = capi.Handle(device=torch.cuda.current_device())
    h.set_stream(stream.cuda_stream)
    Qd = torch.from_numpy(np.ascontiguousarray(Qh)).cuda()
    h.set_q_dense_dev(n3, Qd.data_ptr(), n3)
    R0 = identity_start(torch, n, RANK, "cuda"); s0 = torch.ones(n, dtype=torch.float64, device="cuda")
    R = torch.empty_like(R0); s = torch.empty_like(s0)
    out = {}
    for r in (3, 5, 10):
        X = torch.randn(r, n3, dtype=torch.float64, device="cuda"); O = torch.empty_like(X)
        h.qy_dev(r, X.data_ptr(), O.data_ptr())
        lock = h.bench_qy(r, -50)
        alg = 72.0 * n * n + 48.0 * n * r
        out[f"qy_r{r}"] = {"us_per_product_lockstep": lock * 1e3, "frac_lockstep": alg / (lock * 1e-3) / 1e9 / peak}
    for _ in range(2):
        h.trust_region_dev(RANK, R0.data_ptr(), s0.data_ptr(), R.data_ptr(), s.data_ptr(), lam=LAM, gradtol=GRADTOL)
    ms = its = prods = 0.0
    for _ in range(steps):
        primal, _, st = h.trust_region_dev(RANK, R0.data_ptr(), s0.data_ptr(), R.data_ptr(), s.data_ptr(), lam=LAM, gradtol=GRADTOL)
        ms += st["solve_ms"]; its += st["tcg_iters"]; prods += st["qy_products"]
    alg = 72.0 * n * n + 48.0 * n * RANK
    out.update({"workload": "BAL-Ladybug-1723-shaped synthetic dense Q (213.7 MB), rank-3 full solve from identity to gradnorm<1e-6 (the round-1 bench workload)",
                "value": its / (ms * 1e-3), "unit": "tCG iterations/s", "time_to_kkt_ms": ms / steps, "tcg_iters_per_solve": its / steps,
                "qy_products_per_solve": prods / steps, "final_objective": primal, "exit": st["exit"],
                "roofline_frac": alg * prods / (ms * 1e-3) / 1e9 / peak, "grid_barrier_us": h.bench_barrier(RANK, 2000)})
    # certificate on this problem: the reference's dense route (syevd on 3N x 3N) next to the iterative one, same decision
    Rh = R.cpu().numpy().T.copy(order="F"); sh = s.cpu().numpy()
    cd = h.certify(Rh, sh, LAM, primal, method="dense")
    ci = h.certify(Rh, sh, LAM, primal, method="iterative")
    out["certificate"] = {"dense_syevd_ms": cd["ms"], "iterative_ms": ci["ms"], "iterative_products": ci["products"],
                          "min_eig_dense": cd["min_eig"], "min_eig_iterative": ci["min_eig"], "certified": [cd["certified"], ci["certified"]]}
    h.close()
    return out


def bsr_er100k(torch, capi, xdist, world, rank, local, stream, peak):
    """BASELINE config 5: synthetic Erdos-Renyi view graph, 100k cameras / ~5M edges, block-CSR Q (one 128-byte block per directed edge),
    r in {5, 10, 20}: the block-CSR Q.Y alone and a short time-capped solve, on this run's GPUs (cameras partitioned when world > 1).
    Algorithmic bytes per GPU: nnzb (128 + 4) + 4 (N + 1) + 8 * 3N r (operand) + 8 * 3 rows r (result)."""
    import torch.distributed as dist
    from xm_code_b200 import problems
    N = int(os.environ.get("XM_BENCH_BSR_CAMERAS", "100000"))
    rowptr, col, vals = problems.erdos_renyi_bsr(N, avg_degree=100.0, seed=0)
    h = capi.Handle(device=local)
    lo, hi = 0, N
    if world > 1:
        info = xdist.attach(h, N, 20)
        lo, hi = info["cam_lo"], info["cam_hi"]
    h.set_stream(stream.cuda_stream)
    h.set_q_bsr(rowptr, col, vals, 3)
    nnzb_local = int(rowptr[hi] - rowptr[lo])
    out = {"workload": f"Erdos-Renyi view graph + Hamiltonian path, {N} cameras, {int(rowptr[-1])} directed 3x3 blocks (128 B padded each), block-CSR",
           "nnzb": int(rowptr[-1]), "halo": h.comm_halo() if world > 1 else None, "ranks": {}}
    rng = np.random.default_rng(1)
    for r in (5, 10, 20):
        X = torch.from_numpy(rng.standard_normal((r, 3 * N))).cuda(); O = torch.empty_like(X)
        h.qy_dev(r, X.data_ptr(), O.data_ptr())
        v = [h.bench_qy(r, 20), h.bench_qy(r, -20)]
        R0 = identity_start(torch, N, r, "cuda"); s0 = torch.ones(N, dtype=torch.float64, device="cuda")
        Rd = torch.empty_like(R0); sd = torch.empty_like(s0)
        primal, _, st = h.trust_region_dev(r, R0.data_ptr(), s0.data_ptr(), Rd.data_ptr(), sd.data_ptr(), lam=0.0, gradtol=1e-6, max_time=1.0)
        v += [st["solve_ms"], float(st["tcg_iters"]), float(st["qy_products"])]
        if world > 1:
            t = torch.tensor(v[:3], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX); v[:3] = [float(x) for x in t.tolist()]
        alg = nnzb_local * (128 + 4) + 4 * (hi - lo + 1) + 8 * 3 * N * r + 8 * 3 * (hi - lo) * r
        out["ranks"][f"r{r}"] = {"ms_per_product_free_running": v[0], "ms_per_product_lockstep": v[1], "algorithmic_bytes_per_gpu": alg,
                                 "frac_of_hbm_peak_lockstep": alg / (v[1] * 1e-3) / 1e9 / peak,
                                 "solve_tcg_iters_per_s": v[3] / (v[2] * 1e-3), "solve_ms_per_product": v[2] / max(1.0, v[4])}
    if world > 1:
        xdist.detach(h)
    else:
        h.close()
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist
    from xm_code_b200 import capi, dist as xdist

    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the XM hot path has no CPU fallback (use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local)
    if world > 1:
        import datetime
        dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=datetime.timedelta(minutes=20))
    N = N_CAMERAS; n3 = 3 * N
    t_gen = time.perf_counter()
    Qt, prob = make_problem_device("cuda")                 # every rank builds the same matrix on its own GPU (deterministic)
    torch.cuda.synchronize(); t_gen = time.perf_counter() - t_gen
    h = capi.Handle(device=local, profile=bool(int(os.environ.get("XM_PROFILE", "0"))), qy_variant=int(os.environ.get("XM_QY_VARIANT", "0")),
                    vec_in_global=bool(int(os.environ.get("XM_VEC_GLOBAL", "0"))))
    cam_lo, cam_hi = 0, N
    if world > 1:        # one solve, cameras (rows of Q) partitioned over the ranks; torch.distributed only carries the IPC handles
        info = xdist.attach(h, N, 10)
        cam_lo, cam_hi = info["cam_lo"], info["cam_hi"]
    stream = torch.cuda.current_stream()
    h.set_stream(stream.cuda_stream)
    row0, nrows = 3 * cam_lo, 3 * (cam_hi - cam_lo)
    # this rank's row slab: on the device a view into Qt (column-major, ld = 3N: Q is symmetric); on the host a pinned
    # column-major copy with ld = nrows
    Q_pin = torch.empty((n3, nrows), dtype=torch.float64, pin_memory=True)
    Q_pin.copy_(Qt[:, row0:row0 + nrows])
    slab_dev_ptr = Qt.data_ptr() + 8 * row0

    def upload_q(dev):
        if dev:
            h._check(h.lib.xm_set_q_dense_slab_dev(h._h, n3, row0, nrows, capi.C.c_void_p(slab_dev_ptr), n3), "xm_set_q_dense_slab_dev")
        else:
            h._check(h.lib.xm_set_q_dense_slab(h._h, n3, row0, nrows, capi.C.c_void_p(Q_pin.data_ptr()), nrows), "xm_set_q_dense_slab")
        h.N = N

    upload_q(True)
    torch.cuda.synchronize()
    del Qt
    torch.cuda.empty_cache()
    R0_dev = identity_start(torch, N, RANK, "cuda"); s0_dev = torch.ones(N, dtype=torch.float64, device="cuda")
    R0_pin = R0_dev.cpu().pin_memory(); s0_pin = s0_dev.cpu().pin_memory()
    R_dev = torch.empty_like(R0_dev); s_dev = torch.empty_like(s0_dev)
    R_out = torch.empty_like(R0_pin); s_out = torch.empty_like(s0_pin)
    torch.cuda.synchronize()

    def step_device():
        primal, _, st = h.trust_region_dev(RANK, R0_dev.data_ptr(), s0_dev.data_ptr(), R_dev.data_ptr(), s_dev.data_ptr(),
                                           lam=LAM, gradtol=GRADTOL, max_time=STEP_SECONDS)
        return primal, st

    def step_e2e():
        upload_q(False)
        gt = capi.C.c_double(GRADTOL); pr = capi.C.c_double(); st = capi.XmStats()
        rc = h.lib.xm_trust_region(h._h, RANK, capi.C.c_void_p(R0_pin.data_ptr()), capi.C.c_void_p(s0_pin.data_ptr()), LAM, capi.C.byref(gt),
                                   0.0, None, STEP_SECONDS, capi.C.c_void_p(R_out.data_ptr()), capi.C.c_void_p(s_out.data_ptr()),
                                   capi.C.byref(pr), capi.C.byref(st), None)
        h._check(rc, "xm_trust_region")
        stats = {f[0]: getattr(st, f[0]) for f in capi.XmStats._fields_}
        stats["exit"] = capi.EXIT_CODES.get(st.exit_code, str(st.exit_code)); stats["phase_ms"] = list(st.phase_ms)
        return pr.value, stats

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, K, W):
        for _ in range(W):
            fn()
        barrier()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        acc = {"tcg_iters": 0, "qy_products": 0, "outer_iters": 0, "solve_ms": 0.0, "qy_ms": 0.0, "sync_ms": 0.0}
        last = None
        e0.record(stream)
        for _ in range(K):
            primal, st = fn()
            for k in acc:
                acc[k] += st[k]
            last = (primal, st)
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t.item())
        return ms, acc, last

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    ms_dev, acc_dev, (primal, st) = timed(step_device, args.steps, args.warmup)
    clocks = sampler.stop() if sampler else None
    ms_e2e, acc_e2e, _ = timed(step_e2e, args.steps, max(1, args.warmup // 2) if BIG else max(1, args.warmup // 2))
    # the Q.Y phase alone (collective calls when world > 1): the same device code inside the op-level kernel
    qy_alone = {}
    for r in (3, 5, 10):
        X_dev = torch.randn(r, n3, dtype=torch.float64, device="cuda"); O_dev = torch.empty_like(X_dev)
        h.qy_dev(r, X_dev.data_ptr(), O_dev.data_ptr())
        nrep = 10 if BIG else 50
        v = [h.bench_qy(r, nrep), h.bench_qy(r, -nrep)]
        if world > 1:
            t = torch.tensor(v, device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX); v = [float(x) for x in t.tolist()]
        qy_alone[r] = v
    barrier_us = h.bench_barrier(RANK, 2000)
    # wall-clock-to-KKT: ONE full solve to the tolerance (no time limit)
    full = None
    if BIG and int(os.environ.get("XM_BENCH_FULL_SOLVE", "1")):
        fp, _, fst = h.trust_region_dev(RANK, R0_dev.data_ptr(), s0_dev.data_ptr(), R_dev.data_ptr(), s_dev.data_ptr(), lam=LAM, gradtol=GRADTOL, max_time=1000.0)
        full = {"time_to_kkt_ms": fst["solve_ms"], "tcg_iters": fst["tcg_iters"], "outer_iters": fst["outer_iters"], "qy_products": fst["qy_products"],
                "final_objective": fp, "final_gradnorm": fst["gradnorm"], "exit": fst["exit"], "tcg_iters_per_s": fst["tcg_iters"] / (fst["solve_ms"] * 1e-3)}
    solve_ms = acc_dev["solve_ms"]
    if world > 1:        # the slowest rank's kernel time counts
        t = torch.tensor([solve_ms, barrier_us] + ([full["time_to_kkt_ms"]] if full else []), device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        vals = [float(x) for x in t.tolist()]
        solve_ms, barrier_us = vals[0], vals[1]
        if full:
            full["time_to_kkt_ms"] = vals[2]
    # algorithmic bytes of one Q.Y product ON ONE GPU: its rows of Q once, the operand once, its rows of the result once
    rows_max = 3 * max(hi - lo for lo, hi in xdist.partition_table(N, world, st["grid_ctas"])) if world > 1 else n3
    alg = lambda r: 8.0 * rows_max * n3 + 8.0 * n3 * r + 8.0 * rows_max * r      # noqa: E731
    peak, peak_src = measured_peak_gbs()
    extras = {}
    if world == 1 and rank == 0 and int(os.environ.get("XM_BENCH_EXTRAS", "1")):
        # certificate (f1) on the full-solve point and assembly (f2) of this very problem through the C-ABI
        try:
            if full:
                Rh = R_dev.cpu().numpy().T.copy(order="F"); sh = s_dev.cpu().numpy()
                ci = h.certify(Rh, sh, LAM, full["final_objective"], method="iterative")
                extras["certificate"] = {"method": "block Davidson on the Q.Y operator (xm_certify_ex)", "ms": ci["ms"], "products": ci["products"],
                                         "min_eig": ci["min_eig"], "certified": ci["certified"], "converged": ci["converged"],
                                         "note": "the reference's route is a 41046 x 41046 dense syevd + host LSCG (checkeig.h:190-318)"}
            ha = capi.Handle(device=local)
            ha.set_stream(stream.cuda_stream)
            t0 = time.perf_counter()
            _, _, asm_ms = ha.create_matrix(prob["N"], prob["M"], prob["cam"], prob["lm"], prob["w"], prob["pt"], want_q=False, want_abar=False)
            extras["assembly"] = {"what": "xm_create_matrix: observations -> Q on the device (landmark elimination kernel + Cholesky + SYRK)",
                                  "device_ms": asm_ms, "wall_s": time.perf_counter() - t0, "observations": int(prob["cam"].size), "landmarks": int(prob["M"])}
            ha.close()
        except Exception as e:  # noqa: BLE001
            extras["error"] = f"{type(e).__name__}: {e}"
        extras["latency_regime"] = latency_regime(torch, capi, stream, peak)
    if world > 1:
        xdist.detach(h)
    else:
        h.close()
    if int(os.environ.get("XM_BENCH_BSR", "1")) and int(os.environ.get("XM_BENCH_EXTRAS", "1")):
        torch.cuda.empty_cache()
        try:
            extras["bsr_er100k"] = bsr_er100k(torch, capi, xdist, world, rank, local, stream, peak)
        except Exception as e:  # noqa: BLE001
            extras["bsr_er100k"] = {"error": f"{type(e).__name__}: {e}"}
    if rank != 0:
        return
    # ONE solve shared by all ranks: the job's iterations are the solve's iterations (strong scaling for world > 1)
    value = acc_dev["tcg_iters"] / (ms_dev * 1e-3)
    e2e_value = acc_e2e["tcg_iters"] / (ms_e2e * 1e-3)
    cpu_base = None
    if world == 1:
        cpu_its, cpu_dt, cpu_res, cpu_what, cpu_threads = oracle_sample(Q_pin.numpy(), args.cpu_seconds)      # symmetric: C order == column-major
        cpu_base = {"value": cpu_its, "unit": "tCG iterations/s", "cores": cpu_threads, "kind": "port",
                    "sample": f"{cpu_what}, {cpu_threads} threads of {os.cpu_count()} logical CPUs, the same Q for {cpu_dt:.1f} s "
                              f"({cpu_res.tcg_iters} tCG iterations, {cpu_res.outer_iters} outer)"}
    slab_mb = 8.0 * rows_max * n3 / 1e6
    achieved = alg(RANK) * acc_dev["qy_products"] / (solve_ms * 1e-3) / 1e9
    traffic_pp, traffic_src = measured_traffic(N) if world == 1 else (None, "single-GPU captures only")
    config = dict(CONFIG)
    if slab_mb <= 126:
        config["l2"] = "per-GPU slab of Q = %.1f MB fits the 126 MB L2: after the first product Q.Y streams from L2, the HBM roofline does not bound it" % slab_mb
    line = {
        "metric": "xm_tcg_iterations_per_sec", "value": value, "unit": "tCG iterations/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
        "result": {"tcg_iters_per_step": acc_dev["tcg_iters"] / args.steps, "outer_iters_per_step": acc_dev["outer_iters"] / args.steps,
                   "qy_products_per_step": acc_dev["qy_products"] / args.steps, "objective_at_step_end": primal, "gradnorm_at_step_end": st["gradnorm"],
                   "exit": st["exit"], "full_solve": full, "problem_build_s": t_gen,
                   "grid_ctas": st["grid_ctas"], "threads_per_cta": st["threads_per_cta"], "ksplit": st["ksplit"],
                   "in_kernel_ms_per_step": {"solve": solve_ms / args.steps, "qy": acc_dev["qy_ms"] / args.steps, "grid_sync_wait": acc_dev["sync_ms"] / args.steps},
                   "grid_barrier_us": barrier_us,
                   "parallelism": "single GPU" if world == 1 else
                       f"one solve, cameras partitioned over {world} GPUs (rank 0 owns cameras [{cam_lo},{cam_hi})); per tCG iteration: operand rows "
                       f"+ 2 reduction scalars + 2 barriers exchanged by peer-mapped stores from inside the persistent kernels (no NCCL on the path)",
                   **extras},
        "e2e": {"value": e2e_value, "unit": "tCG iterations/s", "ms_per_step": ms_e2e / args.steps,
                "h2d_bytes_per_step": int(8 * (n3 * n3 + world * (n3 * RANK + N))), "d2h_bytes_per_step": int(world * 8 * (n3 * RANK + N))},
        "gpu_launches": int(args.steps) * world,   # one persistent solve kernel per rank per step (e2e adds the re-layout kernels of the upload)
        # dominant kernel = the persistent solve kernel: one launch executes qy_products dense Q.Y products
        "roofline": {"bound": "hbm", "kernel": "xm_solve_kernel<3,512,0,%s> (persistent: whole XMtrustregion call, Q.Y through the 2-D TMA ring)" % ("false" if world == 1 else "true")
                                               + ("" if world == 1 else " — per GPU, slowest rank"),
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic_pp * acc_dev["qy_products"] / args.steps if traffic_pp else None, "traffic_source": traffic_src,
                     "algorithmic_bytes": alg(RANK) * acc_dev["qy_products"] / args.steps, "algorithmic_bytes_per_product": alg(RANK),
                     "ms_per_launch": solve_ms / args.steps, "products_per_launch": acc_dev["qy_products"] / args.steps, "peak_source": peak_src,
                     "peak_note": "the peak is the pod's measured COPY bandwidth (read + write); this kernel is a read-only stream (ncu: DRAM traffic = "
                                  "1.003 x algorithmic, writes 0.06 %), which HBM3e serves slightly faster than a copy — fractions a few percent above 1.0 are that",
                     "qy_phase_alone": {"kernel": "xm_ops_kernel (same qy_phase device code, MODE_OUT)",
                                        **{f"r{r}": {"us_per_product_free_running": v[0] * 1e3, "us_per_product_lockstep": v[1] * 1e3,
                                                     "frac_lockstep": alg(r) / (v[1] * 1e-3) / 1e9 / peak} for r, v in qy_alone.items()},
                                        "note": "products inside one launch; 'lockstep' adds a grid barrier after every product like the solver; values "
                                                "slightly above 1.0 of the measured COPY peak come from a read-only stream (no write traffic)"}},
        "clocks": clocks,
    }
    if cpu_base:
        line["cpu_baseline"] = cpu_base
    print(json.dumps(line))


def run_reference(args):
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    N = N_CAMERAS; n3 = 3 * N
    harness = os.path.join(ROOT, "oracle", "_ref", "xm_ref_harness")
    have_gpu = False
    try:
        import torch
        have_gpu = torch.cuda.is_available()
    except Exception:
        pass
    base = {"metric": "xm_tcg_iterations_per_sec", "unit": "tCG iterations/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference"}
    if have_gpu and os.path.exists(harness):
        Qt, _ = make_problem_device("cuda")                 # the same generator, the same device type: the same matrix
        Qh = Qt.cpu().numpy()
        del Qt
        torch.cuda.empty_cache()
        need = 8 * n3 * n3 + (1 << 28)
        d = tempfile.mkdtemp(dir="/dev/shm" if shutil.disk_usage("/dev/shm").free > need else None)
        try:
            with open(os.path.join(d, "Q.bin"), "wb") as f:                 # the reference's wire format (XM_main.cu:25-30)
                np.array([n3, n3], dtype=np.int32).tofile(f)
                Qh.tofile(f)                                                # symmetric: C order == column-major
            del Qh
            reps = args.steps + args.warmup
            out = subprocess.run([harness, d, str(RANK), str(GRADTOL), str(LAM), str(STEP_SECONDS), "0", str(reps)], capture_output=True, text=True, timeout=6000)
        finally:
            shutil.rmtree(d, ignore_errors=True)
        js = json.loads(re.search(r"REFJSON (\{.*\})", out.stdout).group(1))
        totals = [int(x) for x in re.findall(r"Total iteration:\s+(\d+)", out.stdout)]
        runs = js["runs"][args.warmup:]; its = totals[args.warmup:]
        tr_ms = sum(r["tr_ms"] for r in runs); e2e_ms = sum(r["h2d_q_ms"] + r["tr_ms"] + r["d2h_ms"] for r in runs)
        value = sum(its) / (tr_ms * 1e-3)                   # device-resident, like our `value`
        e2e_value = sum(its) / (e2e_ms * 1e-3)              # with the H2D of Q and the D2H of R, s, like our `e2e`
        line = dict(base, value=value, ms_per_step=tr_ms / len(runs), config=dict(CONFIG),
                    result={"what": "UNMODIFIED reference XMtrustregion (trustregion.h via oracle/_ref/xm_ref_harness, cuBLAS path) on the same B200; "
                                    "the reference has no CPU implementation of this path",
                            "tcg_iters_per_step": sum(its) / len(its), "objective_at_step_end": runs[-1]["primal"],
                            "h2d_q_ms_per_step": sum(r["h2d_q_ms"] for r in runs) / len(runs)},
                    cpu_baseline={"value": value, "unit": "tCG iterations/s", "cores": 1, "kind": "reference",
                                  "sample": f"{len(runs)} steps; one host thread driving the GPU (reference design)"},
                    e2e={"value": e2e_value, "unit": "tCG iterations/s", "ms_per_step": e2e_ms / len(runs),
                         "h2d_bytes_per_step": int(8 * (n3 * n3 + 3 * n3 + 3 * n3 + 2 * N)), "d2h_bytes_per_step": int(8 * (n3 * RANK + N))})
    else:
        # no GPU / no harness (the build container): the oracle port on the host cores, on a problem the host can assemble
        from xm_code_b200 import problems
        n = min(N, 1723)
        Qh, _ = problems.synthetic_dense_q(n, seed=0, obs_per_camera=OBS_PER_CAMERA, n_landmarks=12 * n)
        its, dt, res, what, threads = oracle_sample(Qh, args.cpu_seconds)
        line = dict(base, value=its, ms_per_step=dt * 1e3, config=dict(CONFIG),
                    result={"what": f"oracle port (restatement of trustregion.h) on the host cores on a {n}-camera problem: the reference harness or a GPU "
                                    "is unavailable; " + what},
                    cpu_baseline={"value": its, "unit": "tCG iterations/s", "cores": threads, "kind": "port",
                                  "sample": f"{what}: {dt:.1f} s of the solve ({res.tcg_iters} tCG iterations)"},
                    e2e={"value": its, "unit": "tCG iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0})
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
