#!/usr/bin/env python
"""bench.py — XM Burer-Monteiro trust-region throughput on B200 (driver contract in the task statement).

Workload: BAL-Ladybug-1723-shaped synthetic dense Q (1723 cameras, 3N = 5169, Q = 213.7 MB FP64 > L2), one
"step" = one full XMtrustregion-equivalent call at rank 3 from the reference's identity start to gradnorm < 1e-6
(reference call: XMtrustregion(C,R0,s0,R,s,lam=0,gradtol=1e-6,ls=0,...), XM/include/XM/trustregion.h:77).
metric = tCG iterations per second (each iteration = one Q.Y + the fused per-camera work), time-to-KKT = ms_per_step.

--gpus N > 1 (torchrun, one process per GPU): the SAME solve partitioned by camera over the N GPUs ("scaling":
"strong") — rank k holds the rows of Q of its cameras; per tCG iteration the persistent kernels exchange the operand
rows and the reduction scalars by peer-mapped stores over NVLink (xm_code_b200/dist.py, include/xm_b200.h).
XM_BENCH_CAMERAS=<n> changes the camera count (e.g. 13682 = BAL-Final-sized dense Q, 13.5 GB).

  value : device-resident (Q, R0, s0 already in HBM; CUDA events on the launch stream)
  e2e   : through the C-ABI with HOST buffers — xm_set_q_dense (pinned H2D of Q + re-layout) + xm_trust_region
          (H2D of R0/s0, solve, D2H of R/s) inside the timed region
  roofline      : the dense Q.Y kernel timed alone (xm_bench_qy), algorithmic bytes 72 N^2 + 48 N r
  cpu_baseline  : the compiled C + OpenMP oracle (port; oracle/xm_oracle_c.c) on all host cores, bounded sample
  --impl reference : the UNMODIFIED reference trustregion.h (oracle/_ref/xm_ref_harness, cuBLAS path) on the same
          Q on the same GPU — the reference has no CPU implementation of this path; falls back to the oracle port
          on the host when the harness or a GPU is missing.
"""
from __future__ import annotations

import argparse
import json
import os
import re
import statistics
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_CAMERAS = int(os.environ.get("XM_BENCH_CAMERAS", "1723"))
RANK = 3
GRADTOL = 1e-6
LAM = 0.0
# ncu, one solve launch on this workload (1684 products): dram read 359,107,278,592 B + write 1,091,153,664 B
NCU_SOLVE_DRAM_BYTES_PER_PRODUCT = (359107278592 + 1091153664) / 1684
WORKLOAD = f"BAL-Ladybug-{N_CAMERAS}-shaped synthetic dense Q (3N={3 * N_CAMERAS}, {72 * N_CAMERAS ** 2 / 1e6:.1f} MB FP64), rank-3 solve from identity to gradnorm<1e-6"


def make_problem():
    from xm_code_b200 import problems
    Q, prob = problems.synthetic_dense_q(N_CAMERAS, seed=0, obs_per_camera=60, n_landmarks=12 * N_CAMERAS)
    return np.asfortranarray(Q), prob


BIG = N_CAMERAS >= 4000      # large-problem mode (e.g. XM_BENCH_CAMERAS=13682, BAL-Final-sized: 13.5 GB of Q)


def make_problem_shared(world, rank):
    """Large problems under torchrun: rank 0 assembles Q once (tens of seconds of host BLAS, ~4x the matrix in host memory)
    and shares it through /dev/shm; every rank maps it read-only and uploads only the rows of its cameras."""
    import torch.distributed as dist
    path = f"/dev/shm/xm_bench_Q_{N_CAMERAS}.npy"
    if rank == 0:
        Q, _ = make_problem()
        np.save(path + ".tmp.npy", np.ascontiguousarray(Q))     # symmetric: C order == column-major
        os.replace(path + ".tmp.npy", path)
        del Q
    dist.barrier()
    return np.load(path, mmap_mode="r")


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
        # large-problem mode: the other ranks wait in a barrier while rank 0 assembles Q on the host (minutes at BAL-Final size)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=datetime.timedelta(minutes=60 if BIG else 10))
    big = BIG and world > 1
    Qh = make_problem_shared(world, rank) if big else make_problem()[0]
    N = N_CAMERAS; n3 = 3 * N
    h = capi.Handle(device=local, profile=bool(int(os.environ.get("XM_PROFILE", "0"))), qy_variant=int(os.environ.get("XM_QY_VARIANT", "0")),
                    vec_in_global=bool(int(os.environ.get("XM_VEC_GLOBAL", "0"))))
    cam_lo, cam_hi = 0, N
    if world > 1:        # one solve, cameras (rows of Q) partitioned over the ranks; torch.distributed only carries the IPC handles
        info = xdist.attach(h, N, RANK)
        cam_lo, cam_hi = info["cam_lo"], info["cam_hi"]
    stream = torch.cuda.current_stream()
    h.set_stream(stream.cuda_stream)
    # host (pinned) and device copies of the inputs, wire layout (column-major)
    if big:      # only this rank's row slab is pinned / resident: column-major slab = the transposed rows, leading dimension nrows
        row0, nrows = 3 * cam_lo, 3 * (cam_hi - cam_lo)
        Q_pin = torch.from_numpy(np.ascontiguousarray(Qh[row0:row0 + nrows, :].T)).pin_memory()
    else:
        Q_pin = torch.from_numpy(Qh.T.copy()).pin_memory()       # memory of the .T copy == column-major Q
    Q_dev = Q_pin.cuda(non_blocking=True)
    R0_np = np.zeros((n3, RANK), order="F")
    for a in range(3):
        R0_np[a::3, a] = 1.0
    R0_pin = torch.from_numpy(np.ascontiguousarray(R0_np.T)).pin_memory(); s0_pin = torch.ones(N, dtype=torch.float64).pin_memory()
    R0_dev = R0_pin.cuda(); s0_dev = s0_pin.cuda()
    R_dev = torch.empty_like(R0_dev); s_dev = torch.empty_like(s0_dev)
    R_out = torch.empty_like(R0_pin); s_out = torch.empty_like(s0_pin)
    torch.cuda.synchronize()

    def step_device():
        primal, _, st = h.trust_region_dev(RANK, R0_dev.data_ptr(), s0_dev.data_ptr(), R_dev.data_ptr(), s_dev.data_ptr(),
                                           lam=LAM, gradtol=GRADTOL)
        return primal, st

    def upload_q(ptr, dev):
        if big:
            fn = h.lib.xm_set_q_dense_slab_dev if dev else h.lib.xm_set_q_dense_slab
            h._check(fn(h._h, n3, row0, nrows, capi.C.c_void_p(ptr), nrows), "xm_set_q_dense_slab")
            h.N = N
        elif dev:
            h.set_q_dense_dev(n3, ptr, n3)
        else:
            h.set_q_dense_ptr(n3, ptr, n3)                      # a rank of a communicator copies only its own rows

    def step_e2e():
        upload_q(Q_pin.data_ptr(), False)
        gt = capi.C.c_double(GRADTOL); pr = capi.C.c_double(); st = capi.XmStats()
        rc = h.lib.xm_trust_region(h._h, RANK, capi.C.c_void_p(R0_pin.data_ptr()), capi.C.c_void_p(s0_pin.data_ptr()), LAM, capi.C.byref(gt),
                                   0.0, None, 1000.0, capi.C.c_void_p(R_out.data_ptr()), capi.C.c_void_p(s_out.data_ptr()),
                                   capi.C.byref(pr), capi.C.byref(st), None)
        h._check(rc, "xm_trust_region")
        return pr.value, st

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, K, W):
        for _ in range(W):
            fn()
        barrier()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        iters = 0; last = None
        e0.record(stream)
        for _ in range(K):
            primal, st = fn()
            iters += st["tcg_iters"] if isinstance(st, dict) else st.tcg_iters
            last = (primal, st)
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t.item())
        return ms, iters, last

    upload_q(Q_dev.data_ptr(), True)
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    ms_dev, it_dev, (primal, st) = timed(step_device, args.steps, args.warmup)
    clocks = sampler.stop() if sampler else None
    ms_e2e, it_e2e, (primal_e, st_e) = timed(step_e2e, args.steps, max(1, args.warmup // 2))
    # roofline of the dominant kernel phase: the dense Q.Y, timed alone (collective calls when world > 1)
    X_dev = torch.randn(RANK, n3, dtype=torch.float64, device="cuda"); O_dev = torch.empty_like(X_dev)
    h.qy_dev(RANK, X_dev.data_ptr(), O_dev.data_ptr())
    qy_ms = h.bench_qy(RANK, 50)                 # 50 products inside one launch, free-running CTAs
    qy_ms_lockstep = h.bench_qy(RANK, -50)       # same with a grid barrier after every product (the solver's regime)
    barrier_us = h.bench_barrier(RANK, 2000)
    solve_ms = st["solve_ms"]
    if world > 1:        # the slowest rank's kernel time counts
        t = torch.tensor([solve_ms, qy_ms, qy_ms_lockstep, barrier_us], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        solve_ms, qy_ms, qy_ms_lockstep, barrier_us = (float(x) for x in t.tolist())
    # algorithmic bytes of one Q.Y product ON ONE GPU: its rows of Q once, the operand once, its rows of the result once
    rows_max = 3 * max(hi - lo for lo, hi in xdist.partition_table(N, world, st["grid_ctas"])) if world > 1 else n3
    alg_bytes = 8.0 * rows_max * n3 + 8.0 * n3 * RANK + 8.0 * rows_max * RANK
    peak, peak_src = measured_peak_gbs()
    if world > 1:
        xdist.detach(h)
    if rank != 0:
        return
    # ONE solve shared by all ranks: the job's iterations are the solve's iterations (strong scaling for world > 1)
    value = it_dev / (ms_dev * 1e-3)
    e2e_value = it_e2e / (ms_e2e * 1e-3)
    cpu_base = None
    if world == 1:
        cpu_its, cpu_dt, cpu_res, cpu_what, cpu_threads = oracle_sample(Qh, args.cpu_seconds)
        cpu_base = {"value": cpu_its, "unit": "tCG iterations/s", "cores": cpu_threads, "kind": "port",
                    "sample": f"{cpu_what}, {cpu_threads} threads of {os.cpu_count()} logical CPUs, the same Q for {cpu_dt:.1f} s "
                              f"({cpu_res.tcg_iters} tCG iterations, {cpu_res.outer_iters} outer)"}
    per_solve = it_dev / args.steps
    slab_mb = 8.0 * rows_max * n3 / 1e6
    achieved = alg_bytes * st["qy_products"] / (solve_ms * 1e-3) / 1e9
    line = {
        "metric": "xm_tcg_iterations_per_sec", "value": value, "unit": "tCG iterations/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak" if world == 1 else "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "cameras": N, "rank": RANK, "gradtol": GRADTOL, "lam": LAM,
                   "tcg_iters_per_solve": per_solve, "outer_iters_per_solve": st["outer_iters"], "qy_products_per_solve": st["qy_products"],
                   "time_to_kkt_ms": ms_dev / args.steps, "final_objective": primal, "final_gradnorm": st["gradnorm"], "exit": st["exit"],
                   "l2": ("inputs larger than L2 (Q = %.1f MB vs 126 MB L2)" % slab_mb) if slab_mb > 126 else
                         ("per-GPU slab of Q = %.1f MB fits the 126 MB L2: after the first product Q.Y streams from L2, the HBM roofline does not bound it" % slab_mb),
                   "grid_ctas": st["grid_ctas"], "threads_per_cta": st["threads_per_cta"], "ksplit": st["ksplit"],
                   "in_kernel_ms": {"solve": solve_ms, "qy": st["qy_ms"], "grid_sync_wait": st["sync_ms"],
                                    "qy_first_tile_wait": st["phase_ms"][0], "qy_tile_waits": st["phase_ms"][1], "qy_tile_math": st["phase_ms"][2],
                                    "qy_reduce_epilogue": st["phase_ms"][3], "profile_timers_on": bool(int(os.environ.get("XM_PROFILE", "0")))},
                   "grid_barrier_us": barrier_us,
                   "parallelism": "single GPU" if world == 1 else
                       f"one solve, cameras partitioned over {world} GPUs (rank 0 owns cameras [{cam_lo},{cam_hi})); per tCG iteration: operand rows "
                       f"+ 2 reduction scalars + 3 barriers exchanged by peer-mapped stores from inside the persistent kernels (no NCCL on the path)"},
        "e2e": {"value": e2e_value, "unit": "tCG iterations/s", "ms_per_step": ms_e2e / args.steps,
                "h2d_bytes_per_step": int(8 * (n3 * n3 + world * (n3 * RANK + N))), "d2h_bytes_per_step": int(world * 8 * (n3 * RANK + N))},
        "gpu_launches": int(args.steps) * world,   # one persistent solve kernel per rank per step (e2e adds one re-layout kernel per step)
        # dominant kernel = the persistent solve kernel (96.6 % of the step in profiles/r01_ncu_solve_and_launches.txt): one launch
        # executes qy_products dense Q.Y products; algorithmic bytes per product (per GPU) = 8 rows 3N + 8 (3N + rows) r
        "roofline": {"bound": "hbm", "kernel": "xm_solve_kernel<3,512,0> (persistent: whole XMtrustregion call, Q.Y through the 2-D TMA ring)"
                                               + ("" if world == 1 else " — per GPU, slowest rank"),
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": NCU_SOLVE_DRAM_BYTES_PER_PRODUCT * st["qy_products"] if (world == 1 and N == 1723) else None,
                     "traffic_source": "ncu dram__bytes_read.sum + dram__bytes_write.sum of one solve launch (profiles/r01_ncu_solve_and_launches.txt), per product",
                     "algorithmic_bytes": alg_bytes * st["qy_products"], "algorithmic_bytes_per_product": alg_bytes,
                     "ms_per_launch": solve_ms, "products_per_launch": st["qy_products"], "peak_source": peak_src,
                     "qy_phase_alone": {"kernel": "xm_ops_kernel<3,512,0> (same qy_phase device code, MODE_OUT)",
                                        "us_per_product_free_running": qy_ms * 1e3, "us_per_product_lockstep": qy_ms_lockstep * 1e3,
                                        "achieved_lockstep": alg_bytes / (qy_ms_lockstep * 1e-3) / 1e9,
                                        "frac_lockstep": alg_bytes / (qy_ms_lockstep * 1e-3) / 1e9 / peak,
                                        "note": "50 products inside one launch; 'lockstep' adds a grid barrier after every product like the solver; "
                                                "values above 1.0 of the measured copy peak come from read-only streaming plus L2 hits on the re-read Q; "
                                                "a single cold product under ncu: 38.9 us, dram read 214.56 MB (profiles/r01_qy_tma_full.summary.txt)"}},
        "clocks": clocks,
    }
    if cpu_base:
        line["cpu_baseline"] = cpu_base
    print(json.dumps(line))


def run_reference(args):
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    Qh, _ = make_problem()
    N = N_CAMERAS; n3 = 3 * N
    harness = os.path.join(ROOT, "oracle", "_ref", "xm_ref_harness")
    have_gpu = False
    try:
        import torch
        have_gpu = torch.cuda.is_available()
    except Exception:
        pass
    base = {"metric": "xm_tcg_iterations_per_sec", "unit": "tCG iterations/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "higher_is_better": True, "scaling": "weak" if world == 1 else "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference"}
    if have_gpu and os.path.exists(harness):
        from xm_code_b200 import binio
        d = tempfile.mkdtemp()
        binio.save_matrix_to_bin(os.path.join(d, "Q.bin"), Qh)
        reps = args.steps + args.warmup
        out = subprocess.run([harness, d, str(RANK), str(GRADTOL), str(LAM), "1000", "0", str(reps)], capture_output=True, text=True, timeout=3000)
        js = json.loads(re.search(r"REFJSON (\{.*\})", out.stdout).group(1))
        totals = [int(x) for x in re.findall(r"Total iteration:\s+(\d+)", out.stdout)]
        runs = js["runs"][args.warmup:]; its = totals[args.warmup:]
        tr_ms = sum(r["tr_ms"] for r in runs); e2e_ms = sum(r["h2d_q_ms"] + r["tr_ms"] + r["d2h_ms"] for r in runs)
        value = sum(its) / (e2e_ms * 1e-3)
        line = dict(base, value=value, ms_per_step=e2e_ms / len(runs),
                    config={"workload": WORKLOAD, "cameras": N, "rank": RANK, "gradtol": GRADTOL, "lam": LAM,
                            "what": "UNMODIFIED reference XMtrustregion (trustregion.h via oracle/_ref/xm_ref_harness, cuBLAS path) on the same B200; "
                                    "the reference has no CPU implementation of this path",
                            "tcg_iters_per_solve": sum(its) / len(its), "device_only_value": sum(its) / (tr_ms * 1e-3), "final_objective": runs[-1]["primal"]},
                    cpu_baseline={"value": value, "unit": "tCG iterations/s", "cores": 1, "kind": "reference",
                                  "sample": f"{len(runs)} full solves; one host thread driving the GPU (reference design)"},
                    e2e={"value": value, "unit": "tCG iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0})
    else:
        its, dt, res, what, threads = oracle_sample(Qh, args.cpu_seconds)
        line = dict(base, value=its, ms_per_step=dt * 1e3,
                    config={"workload": WORKLOAD, "cameras": N, "rank": RANK, "gradtol": GRADTOL, "lam": LAM,
                            "what": "oracle port (restatement of trustregion.h) on the host cores: the reference harness or a GPU is unavailable; " + what},
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
