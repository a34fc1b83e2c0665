"""compute-sanitizer target (SURVEY.md §5 "race detection"): one small pass over every kernel family of the library —
dense Q.Y (TMA ring + direct), block-CSR Q.Y, the op-level kernels, a SIMPLE2 solve, the certificate (dense + iterative), the
assembly, the recovery, and a world-2 LOOP-BACK team (two members on one GPU: barrier_multi, st_operand / unpack_operand, the
plain-push protocol).  Run as
    XM_WATCHDOG_SCALE=200 compute-sanitizer --tool memcheck  python tools/sanitize_target.py
    XM_WATCHDOG_SCALE=200 compute-sanitizer --tool racecheck python tools/sanitize_target.py
    XM_WATCHDOG_SCALE=200 compute-sanitizer --tool synccheck python tools/sanitize_target.py
(the watchdog scale keeps the in-kernel barrier deadlines from firing under the tools' 10-100x slow-down)."""
import os
import sys
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
from xm_code_b200 import capi, problems  # noqa: E402
from oracle import xm_oracle as xo  # noqa: E402

WHAT = set((sys.argv[1] if len(sys.argv) > 1 else "single,multi").split(","))
rng = np.random.default_rng(0)


def rel(a, b):
    return float(np.max(np.abs(a - b)) / max(1e-300, np.max(np.abs(b))))


if "single" in WHAT:
    Q2 = np.load(os.path.join(ROOT, "tests", "golden", "simple2_Q_ref.npz"))["Q"]
    N = Q2.shape[0] // 3
    for kw in (dict(), dict(qy_variant=1), dict(vec_in_global=True), dict(three_barrier_tcg=True)):
        h = capi.Handle(device=0, **kw)
        h.set_q_dense(Q2)
        for r in (3, 5, 8, 10, 12):          # 8 / 10: two cameras per consumer warp (256-thread CTAs)
            X = rng.standard_normal((3 * N, r))
            assert rel(h.qy(X), Q2 @ X) < 1e-12
        got = h.trust_region(xo.from_blocks(xo.identity_init(N, 3)), np.ones(N), 0.0, 1e-6)
        print("solve", kw, got.primal, got.stats["tcg_iters"], flush=True)
        assert abs(got.primal - 4.8322430007e-02) < 1e-6
        h.close()
    h = capi.Handle(device=0)
    h.set_q_dense(Q2)
    Y10 = xo.mgs_rows(rng.standard_normal((N, 3, 10)))
    g10 = h.trust_region(xo.from_blocks(Y10), np.ones(N), 0.0, 1e-6)
    print("solve r=10", g10.primal, g10.stats["tcg_iters"], g10.stats["threads_per_cta"], flush=True)
    Y = xo.mgs_rows(rng.standard_normal((N, 3, 4))); s = np.concatenate([[1.0], rng.uniform(0.8, 1.2, N - 1)])
    R = xo.from_blocks(Y)
    h.objective(R, s, 0.1); h.rgrad(R, s, 0.1)
    h.rhess(R, s, xo.from_blocks(rng.standard_normal(Y.shape)), rng.standard_normal(N), 0.1)
    h.retract(R, s, xo.from_blocks(0.1 * rng.standard_normal(Y.shape)), 0.1 * rng.standard_normal(N), 0.5)
    got = h.trust_region(xo.from_blocks(xo.identity_init(N, 3)), np.ones(N), 0.0, 1e-8)
    for m in ("dense", "iterative"):
        c = h.certify(got.R, got.s, 0.0, got.primal, method=m)
        print("certify", m, c["certified"], c["min_eig"], c["products"], flush=True)
    out = h.solve(4, 1e-6, 0.0)
    print("staircase", out["rank"], out["status"], flush=True)
    rec = h.recover(got.R, got.s)
    rowptr, col, vals = problems.erdos_renyi_bsr(90, avg_degree=7, seed=1)
    h.set_q_bsr(rowptr, col, vals, 3)
    Qb = problems.bsr_to_dense(rowptr, col, vals)
    for r in (3, 10, 20):
        X = rng.standard_normal((270, r))
        assert rel(h.qy(X), Qb @ X) < 1e-12
    gb = h.trust_region(xo.from_blocks(xo.identity_init(90, 4)), np.ones(90), 0.0, 1e-7)
    print("bsr solve", gb.primal, gb.stats["tcg_iters"], flush=True)
    prob = problems.synthetic_sfm(20, n_landmarks=120, obs_per_camera=30, seed=3)
    Qa, Ab, _ = h.create_matrix(prob["N"], prob["M"], prob["cam"], prob["lm"], prob["w"], prob["pt"], want_q=True, want_abar=True)
    Qh = problems.q_from_observations(prob["N"], prob["M"], prob["cam"], prob["lm"], prob["w"], prob["pt"])
    assert rel(Qa, Qh) < 1e-10
    h.close()
    print("single ok", flush=True)

if "multi" in WHAT:
    world = 2
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    for case in ("dense", "bsr-plain", "bsr-halo"):
        if case == "dense":
            Nn = 96; Qm, _ = problems.synthetic_dense_q(Nn, seed=2)
        elif case == "bsr-plain":
            Nn = 120; rp, cc, vv = problems.erdos_renyi_bsr(Nn, avg_degree=6, seed=2); Qm = problems.bsr_to_dense(rp, cc, vv)
            os.environ["XM_TUNE_PUSH"] = "1"
        else:
            Nn = 160; rp, cc, vv = problems.banded_bsr(Nn, 4, seed=2); Qm = problems.bsr_to_dense(rp, cc, vv)
            os.environ.pop("XM_TUNE_PUSH", None)
        hs = [capi.Handle(device=0, grid_ctas=sms // world) for _ in range(world)]
        streams = [torch.cuda.Stream(device=0) for _ in range(world)]
        for k, (hh, st) in enumerate(zip(hs, streams)):
            hh.set_stream(st.cuda_stream)
            hh.comm_init(k, world, Nn, 5)
        ptrs = [hh.comm_arena() for hh in hs]
        for hh in hs:
            hh.comm_connect_ptrs(ptrs)
        pool = ThreadPoolExecutor(world)

        def call(name, *a, **k):
            return [f.result(timeout=1200) for f in [pool.submit(getattr(hh, name), *a, **k) for hh in hs]]
        if case == "dense":
            call("set_q_dense", Qm)
        else:
            call("set_q_bsr", rp, cc, vv, 3)
        X = rng.standard_normal((3 * Nn, 4))
        for o in call("qy", X, 1.0):
            assert rel(o, Qm @ X) < 1e-12
        tol = 1.0 if case == "bsr-halo" else 1e-7           # a chain-like graph needs thousands of iterations to 1e-7: stop early, same path
        ref = xo.trust_region(Qm, xo.identity_init(Nn, 4), np.ones(Nn), 0.0, tol)
        for g in call("trust_region", xo.from_blocks(xo.identity_init(Nn, 4)), np.ones(Nn), 0.0, tol):
            assert abs(g.primal - ref.primal) <= 1e-7 * abs(ref.primal), (case, g.primal, ref.primal)
        print("multi", case, "ok", flush=True)
        for hh in hs:
            hh.comm_disconnect()
        for hh in hs:
            hh.close()
    print("multi ok", flush=True)
