"""GPU parity, solver level: the persistent trust-region kernel against the oracle (which is itself pinned against
the reference's own output, tests/test_oracle_vs_reference.py) on the reference's shipped inputs, plus the pybind11
module end to end."""
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import xm_oracle as xo
from conftest import anchored_gram, ROOT

pytestmark = pytest.mark.gpu


def check_point(got, ref, primal_rel=1e-10, s_abs=1e-8, x_abs=1e-7):
    assert abs(got.primal - ref.primal) <= primal_rel * abs(ref.primal)
    np.testing.assert_allclose(got.s, ref.s, atol=s_abs, rtol=0)
    np.testing.assert_allclose(anchored_gram(got.R, got.s), anchored_gram(xo.from_blocks(ref.Y), ref.s), atol=x_abs, rtol=0)


def test_simple1_rank3_matches_oracle(gpu_handle_factory, simple1_q):
    N = simple1_q.shape[0] // 3
    h = gpu_handle_factory()
    h.set_q_dense(simple1_q)
    got = h.trust_region(xo.from_blocks(xo.identity_init(N, 3)), np.ones(N), 0.0, 1e-16)
    ref = xo.trust_region(simple1_q, xo.identity_init(N, 3), np.ones(N), 0.0, 1e-16)
    assert abs(got.primal - 2.550991567720) < 5e-11          # certified global optimum
    check_point(got, ref)
    # same trajectory: identical outer-iteration table while the problem is well conditioned
    assert got.stats["outer_iters"] == ref.outer_iters == 13
    for n, (a, b) in enumerate(zip(got.log[:11], ref.log[:11])):
        assert a[0] == b[0] and a[1] == b[1] and a[4] == b[4] and a[5] == b[5]
        # rounding differences grow along the trajectory: tight for the first iterations, looser near convergence
        assert abs(a[2] - b[2]) <= (1e-10 if n < 8 else 1e-8) * abs(b[2])
        assert abs(a[3] - b[3]) <= (1e-8 if n < 8 else 5e-2) * abs(b[3])
    assert got.stats["exit"] == "rdotr_tiny" and got.gradtol == 1e-16
    assert got.stats["qy_products"] > got.stats["tcg_iters"] - got.stats["outer_iters"]


def test_simple2_matches_oracle(gpu_handle_factory, simple2_q):
    N = simple2_q.shape[0] // 3
    h = gpu_handle_factory()
    h.set_q_dense(simple2_q)
    for tol in (1e-1, 1e-10):
        got = h.trust_region(xo.from_blocks(xo.identity_init(N, 3)), np.ones(N), 0.0, tol)
        ref = xo.trust_region(simple2_q, xo.identity_init(N, 3), np.ones(N), 0.0, tol)
        assert got.stats["outer_iters"] == ref.outer_iters
        check_point(got, ref, primal_rel=1e-9 if tol > 1e-5 else 1e-10, s_abs=1e-7 if tol > 1e-5 else 1e-8)
        assert got.gradtol == pytest.approx(ref.gradtol)     # quirk Q1: /10 only on a small-gradient exit
        assert got.stats["exit"] in ("gradtol", "rdotr_tiny")
    assert abs(got.primal - 4.8322430007e-02) < 1e-10


def test_regulariser_and_geometries(gpu_handle_factory):
    from xm_code_b200 import problems
    Q, prob = problems.synthetic_dense_q(80, seed=5)
    N = 80
    ref = xo.trust_region(Q, xo.identity_init(N, 3), np.ones(N), 0.05, 1e-8)
    for kw in (dict(), dict(grid_ctas=1), dict(grid_ctas=7, ksplit=2), dict(grid_ctas=80, ksplit=16)):
        h = gpu_handle_factory(**kw)
        h.set_q_dense(Q)
        got = h.trust_region(xo.from_blocks(xo.identity_init(N, 3)), np.ones(N), 0.05, 1e-8)
        check_point(got, ref, primal_rel=1e-9, s_abs=1e-7, x_abs=1e-6)


def test_rank_escalation_line_search_path(gpu_handle_factory):
    """Replays an r -> r+1 escalation (XM_main.cu:265-271 + trustregion.h:360-408) with identical inputs on both sides,
    including quirk Q3 (stale sR after an accepted line search)."""
    rng = np.random.default_rng(11)
    N = 30
    A = rng.standard_normal((3 * N, 3 * N + 2))
    Q = A @ A.T / (3 * N)                     # generic PSD matrix: rank 3 is not tight, escalation does real work
    h = gpu_handle_factory()
    h.set_q_dense(Q)
    res3 = xo.trust_region(Q, xo.identity_init(N, 3), np.ones(N), 0.0, 1e-7)
    c = xo.certificate(Q, xo.from_blocks(res3.Y * res3.s[:, None, None]), 0.0, res3.primal)
    assert not c["certified"]
    Y0 = np.concatenate([res3.Y, np.zeros((N, 3, 1))], axis=2)
    v = (c["v"].reshape(N, 3) / res3.s[:, None]).reshape(-1)
    ref = xo.trust_region(Q, Y0, res3.s, 0.0, 1e-7, ls_step=1.0, v=v)
    got = h.trust_region(xo.from_blocks(Y0), res3.s, 0.0, 1e-7, ls_step=1.0, v=v)
    assert ref.status == 0 and ref.primal < res3.primal
    assert got.log[0][2] == pytest.approx(ref.log[0][2], rel=1e-12)      # loss[0] is the stale f0 (Q3)
    assert abs(got.primal - ref.primal) <= 1e-8 * abs(ref.primal)
    # a direction whose trial objective is not below f0 (NaN compares false both ways) -> line search failure,
    # primal = -1 (trustregion.h:394-405); deterministic on both sides
    Yopt = np.concatenate([ref.Y, np.zeros((N, 3, 1))], axis=2)
    bad = h.trust_region(xo.from_blocks(Yopt), ref.s, 0.0, 1e-7, ls_step=1.0, v=np.full(3 * N, np.nan))
    assert bad.primal == -1.0 and bad.stats["exit"] == "linesearch_failed"
    bad_ref = xo.trust_region(Q, Yopt, ref.s, 0.0, 1e-7, ls_step=1.0, v=np.full(3 * N, np.nan))
    assert bad_ref.primal == -1.0


def test_certificate_matches_oracle(gpu_handle_factory, simple1_q):
    N = simple1_q.shape[0] // 3
    h = gpu_handle_factory()
    h.set_q_dense(simple1_q)
    got = h.trust_region(xo.from_blocks(xo.identity_init(N, 3)), np.ones(N), 0.0, 1e-16)
    c = h.certify(got.R, got.s, 0.0, got.primal)
    ref = xo.certificate(simple1_q, got.R * np.repeat(got.s, 3)[:, None], 0.0, got.primal)
    assert c["certified"] and ref["certified"]
    assert abs(c["min_eig"] - ref["min_eig"]) < 1e-8 and abs(c["dual"] - ref["dual"]) < 1e-7
    # a non-tight case: the eigenvector must be a descent direction (sign-free) with the same eigenvalue
    rng = np.random.default_rng(12)
    A = rng.standard_normal((60, 62)); Q = A @ A.T / 60
    h.set_q_dense(Q)
    r3 = h.trust_region(xo.from_blocks(xo.identity_init(20, 3)), np.ones(20), 0.0, 1e-8)
    c = h.certify(r3.R, r3.s, 0.0, r3.primal)
    ref = xo.certificate(Q, r3.R * np.repeat(r3.s, 3)[:, None], 0.0, r3.primal)
    assert c["certified"] == ref["certified"]
    assert abs(c["min_eig"] - ref["min_eig"]) < 1e-8 * max(1.0, abs(ref["min_eig"]))
    assert min(np.abs(c["v"] - ref["v"]).max(), np.abs(c["v"] + ref["v"]).max()) < 1e-6


def test_xm_module_runs_the_reference_demo_call(tmp_path, simple1_q):
    """What 1_test_solve.py does (XM.solve(path, 3, 1e-16, 0.0, 1000)), through the compiled pybind11 module."""
    from xm_code_b200 import binio
    d = tmp_path / "SIMPLE1"
    d.mkdir()
    binio.save_matrix_to_bin(str(d / "Q.bin"), simple1_q)
    code = ("import sys; sys.path.append(%r); import XM; XM.solve(%r, 3, 1e-16, 0.0, 1000); "
            "print('status', XM.solve_rebuttle(%r, 4, 1e-6, 0.0, 1000)); XM.solve_rank3(%r, 3, 1e-6, 0.0, 1000)") % (
        os.path.join(ROOT, "XM", "build"), str(d) + "/", str(d), str(d))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "BM finished with rank 3" in out.stdout and "status 1" in out.stdout
    R = binio.load_matrix_from_bin(str(d / "R.bin")); s = binio.load_matrix_from_bin(str(d / "s.bin"))
    assert R.shape == (447, 3) and s.shape == (149, 1) and s[0, 0] == 1.0
    ref = xo.solve(simple1_q, 3, 1e-6, 0.0, rank3_only=True)
    np.testing.assert_allclose(s[:, 0], ref["s"], atol=1e-6)


def test_full_size_properties_bal_shaped(gpu_handle_factory):
    """BASELINE-size check through size-independent properties (the oracle would take minutes here):
    feasibility of the returned point, monotone accepted losses, optimality (small Riemannian gradient via the op
    hook), agreement of the objective with an independent NumPy evaluation, and run-to-run determinism."""
    from xm_code_b200 import problems
    Q, prob = problems.synthetic_dense_q(1723, seed=0, obs_per_camera=60, n_landmarks=12 * 1723)
    N = 1723
    h = gpu_handle_factory()
    h.set_q_dense(Q)
    R0 = xo.from_blocks(xo.identity_init(N, 3))
    a = h.trust_region(R0, np.ones(N), 0.0, 1e-6)
    b = h.trust_region(R0, np.ones(N), 0.0, 1e-6)
    assert a.stats["exit"] in ("gradtol", "rdotr_tiny")
    assert np.array_equal(a.R, b.R) and np.array_equal(a.s, b.s) and a.primal == b.primal      # deterministic
    Y = xo.to_blocks(a.R)
    np.testing.assert_allclose(np.einsum("iaj,ibj->iab", Y, Y), np.broadcast_to(np.eye(3), (N, 3, 3)), atol=1e-12)
    sR = a.R * np.repeat(a.s, 3)[:, None]
    assert abs(float(np.vdot(Q @ sR, sR)) - a.primal) <= 1e-10 * abs(a.primal)
    losses = [l[2] for l in a.log]
    assert all(x >= y - 1e-12 * abs(x) for x, y in zip(losses, losses[1:]))
    _, _, gn = h.rgrad(a.R, a.s, 0.0)
    assert gn < 1e-6
    assert np.max(np.abs(a.s - prob["s"])) < 0.15               # recovers the ground-truth scales up to noise


def test_simple2_pipeline_end_to_end(tmp_path, simple2_obs, gpu_handle_factory):
    """BASELINE config 2 — what 2_test_creatematrix.py does: observations -> create_matrix -> Q.bin / Abar.bin ->
    XM.solve(path, 5, 1e-1, 0, 1000) (compiled module) -> R.bin / s.bin -> recover_XM (GPU) -> rotations against the shipped
    ground truth gtR (relative to camera 1; SURVEY.md §8c: median 0.09 deg, max 0.9 deg at this tolerance)."""
    from xm_code_b200 import binio, creatematrix
    from xm_code_b200.recover import recover_XM
    o = simple2_obs
    N, M = int(o["N"]), int(o["M"])
    d = tmp_path / "SIMPLE2"
    d.mkdir()
    creatematrix.create_matrix(o["weights"], o["edges"], o["pts"], str(d))
    code = "import sys; sys.path.append(%r); import XM; XM.solve(%r, 5, 1e-1, 0.0, 1000)" % (os.path.join(ROOT, "XM", "build"), str(d))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    R = binio.load_matrix_from_bin(str(d / "R.bin")); s = binio.load_matrix_from_bin(str(d / "s.bin"))
    Q = binio.load_matrix_from_bin(str(d / "Q.bin")); Abar = binio.load_matrix_from_bin(str(d / "Abar.bin"))
    assert R.shape[0] == 3 * N and R.shape[1] in (3, 4) and s.shape == (N, 1)        # certified at rank 3 or 4
    R_real, s_real, p_est, t_est = recover_XM(Q, R, s, Abar, 0.0, handle=gpu_handle_factory())
    assert R_real.shape == (3, 3 * N) and p_est.shape == (3, M) and t_est.shape == (3, N)
    orig = np.load(os.path.join(ROOT, "tests", "golden", "simple2_frames.npz"))["orig_of_new"]
    G = o["gtR"].reshape(3, -1, 3).transpose(1, 0, 2)[orig]
    Rb = R_real.reshape(3, N, 3).transpose(1, 0, 2)
    err = []
    for i in range(N):
        c = (np.trace(Rb[i].T @ (G[0] @ G[i].T)) - 1.0) / 2.0
        err.append(np.degrees(np.arccos(np.clip(c, -1.0, 1.0))))
    assert np.median(err) < 0.2 and np.max(err) < 1.5, (np.median(err), np.max(err))
    assert abs(np.mean(s_real) - 1.0) < 0.01 and np.all(s_real > 0.9) and np.all(s_real < 1.1)
    # and the recovery agrees with the oracle's restatement of recover_XM on the same solver output
    ref = xo.recover(R, s[:, 0], Abar)
    np.testing.assert_allclose(R_real, ref["R"], atol=1e-10)
    np.testing.assert_allclose(t_est, ref["t"], atol=1e-8 * max(1.0, np.abs(ref["t"]).max()))


@pytest.mark.parametrize("erec", ["1"])
def test_two_barrier_iteration_keeps_parity(gpu_handle_factory, simple1_q, simple2_q, erec, monkeypatch):
    """EXPERIMENT (branch tcg2): XM_TUNE_EREC=1 — E = 2QX(p) by recurrence, operand from the new residual, two barriers per
    tCG iteration.  Same parity bars as the standard iteration (oracle study: tests/test_oracle.py)."""
    monkeypatch.setenv("XM_TUNE_EREC", erec)
    from xm_code_b200 import problems
    for Q, r, lam, tol in ((simple1_q, 3, 0.0, 1e-16), (simple2_q, 3, 0.0, 1e-10), (problems.synthetic_dense_q(120, seed=5)[0], 5, 0.05, 1e-8)):
        N = Q.shape[0] // 3
        h = gpu_handle_factory()
        h.set_q_dense(Q)
        Y0 = xo.identity_init(N, r)
        got = h.trust_region(xo.from_blocks(Y0), np.ones(N), lam, tol)
        ref = xo.trust_region(Q, Y0, np.ones(N), lam, tol)
        check_point(got, ref, primal_rel=1e-9, s_abs=1e-7, x_abs=1e-6)
        assert got.stats["outer_iters"] == ref.outer_iters
