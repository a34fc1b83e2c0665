#!/usr/bin/env python
"""Block-CSR Q.Y at Erdos-Renyi scale (BASELINE config 5: 100k cameras / 5M edges, r in {5,10,20}) through the C-ABI.

  python tools/bench_bsr.py [--cameras 100000] [--degree 100] [--ranks 5,10,20] [--solve]
  torchrun --nproc-per-node W tools/bench_bsr.py ...      (cameras partitioned over W GPUs, xm_code_b200/dist.py)

Prints one JSON line per rank r: device time per Q.Y product (CUDA events, 20 products inside one launch, with and
without a barrier after each), algorithmic bytes (nnzb (128 + 4) + 4 (N + 1) + 2 * 8 * 3N r with the
128-byte padded blocks this library stores) and the fraction of the measured HBM peak.  --solve adds a short
trust-region run (a few outer iterations) for tCG iterations/s at that scale.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cameras", type=int, default=100000)
    ap.add_argument("--degree", type=float, default=100.0)
    ap.add_argument("--ranks", default="5,10,20")
    ap.add_argument("--solve", action="store_true")
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    import torch
    from xm_code_b200 import capi, problems, dist as xdist

    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    t0 = time.time()
    rowptr, col, vals = problems.erdos_renyi_bsr(args.cameras, avg_degree=args.degree, seed=0)
    N = args.cameras; nnzb = int(rowptr[-1])
    ranks = [int(x) for x in args.ranks.split(",")]
    h = capi.Handle(device=local)
    if world > 1:
        xdist.attach(h, N, max(ranks))
    h.set_q_bsr(rowptr, col, vals, 3)
    t_setup = time.time() - t0
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        peak = 6650.0
    lo, hi = (0, N)
    if world > 1:
        info = h.comm_info(); lo, hi = info["cam_lo"], info["cam_hi"]
    nnzb_local = int(rowptr[hi] - rowptr[lo])
    lines = []
    rng = np.random.default_rng(1)
    for r in ranks:
        X = torch.from_numpy(rng.standard_normal((r, 3 * N))).cuda()       # memory == 3N x r column-major
        O = torch.empty_like(X)
        h.qy_dev(r, X.data_ptr(), O.data_ptr())
        if r == ranks[0] and N <= 20000 and world == 1:                      # spot check against SciPy on small runs
            import scipy.sparse as sp
            A = sp.bsr_matrix((np.swapaxes(vals, 1, 2), col, rowptr), shape=(3 * N, 3 * N))
            ref = A @ X.cpu().numpy().T
            err = np.max(np.abs(O.cpu().numpy().T - ref)) / np.max(np.abs(ref))
            assert err < 1e-12, err
        free_ms = h.bench_qy(r, 20)
        lock_ms = h.bench_qy(r, -20)
        if world > 1:
            t = torch.tensor([free_ms, lock_ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX); free_ms, lock_ms = (float(x) for x in t.tolist())
        # per GPU: its block rows (values + column indices + row pointers), the whole operand, its rows of the result
        alg = nnzb_local * (128 + 4) + 4 * (hi - lo + 1) + 8 * 3 * N * r + 8 * 3 * (hi - lo) * r
        line = {"what": "bsr_qy", "cameras": N, "nnzb": nnzb, "n_gpus": world, "rank_r": r, "nnzb_per_gpu": nnzb_local,
                "ms_per_product_free_running": free_ms, "ms_per_product_lockstep": lock_ms,
                "algorithmic_bytes_per_gpu": alg, "achieved_gbs_per_gpu": alg / (free_ms * 1e-3) / 1e9,
                "achieved_gbs_per_gpu_lockstep": alg / (lock_ms * 1e-3) / 1e9, "peak_gbs": peak,
                "frac": alg / (free_ms * 1e-3) / 1e9 / peak, "frac_lockstep": alg / (lock_ms * 1e-3) / 1e9 / peak,
                "flops": 2 * 9 * r * nnzb, "setup_s": t_setup}
        if args.solve:
            hs = h
            R0 = np.zeros((3 * N, r), order="F")
            for a in range(3):
                R0[a::3, a] = 1.0
            R0d = torch.from_numpy(np.ascontiguousarray(R0.T)).cuda(); s0d = torch.ones(N, dtype=torch.float64, device="cuda")
            Rd = torch.empty_like(R0d); sd = torch.empty_like(s0d)
            # bounded: stop on the time limit (the reference's max_time exit, trustregion.h:538-543)
            primal, _, st = hs.trust_region_dev(r, R0d.data_ptr(), s0d.data_ptr(), Rd.data_ptr(), sd.data_ptr(), lam=0.0, gradtol=1e-6, max_time=2.0)
            line["solve"] = {"tcg_iters": st["tcg_iters"], "outer_iters": st["outer_iters"], "qy_products": st["qy_products"], "solve_ms": st["solve_ms"],
                             "tcg_iters_per_s": st["tcg_iters"] / (st["solve_ms"] * 1e-3), "exit": st["exit"], "primal": primal,
                             "ms_per_qy_product_in_solve": st["solve_ms"] / max(1, st["qy_products"])}
        lines.append(line)
        if rank == 0:
            print(json.dumps(line), flush=True)
    if world > 1:
        xdist.detach(h)
    if rank == 0 and args.out:
        with open(args.out, "w") as f:
            for l in lines:
                f.write(json.dumps(l) + "\n")


if __name__ == "__main__":
    main()
