"""Matrix-free certificate and the Python staircase over the C-ABI.

CPU: `certificate.certify` with a NumPy operator against the oracle's dense restatement of checkeig (global least squares +
full `eigh`), and `solver.solve_arrays` driven by an oracle-backed stand-in for the GPU handle against the oracle's own
staircase — this checks the control flow (escalation, zero-padded column, v / s scaling, gradtol carry-over, statuses).
GPU: the same through the real handle."""
import numpy as np
import pytest

from oracle import xm_oracle as xo
from xm_code_b200 import certificate, problems, solver


def nontight(N=30, seed=11):
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((3 * N, 3 * N + 2))
    return A @ A.T / (3 * N)


def point(Q, r, lam, tol):
    N = Q.shape[0] // 3
    res = xo.trust_region(Q, xo.identity_init(N, r), np.ones(N), lam, tol)
    return xo.from_blocks(res.Y), res.s, res.primal


CASES = [("simple1", 3, 0.0, 1e-16), ("simple2", 3, 0.0, 1e-1), ("simple2", 3, 0.0, 1e-10), ("nontight", 3, 0.0, 1e-7),
         ("syn", 3, 0.3, 1e-6), ("syn", 5, 0.05, 1e-7)]


@pytest.mark.parametrize("name,r,lam,tol", CASES)
def test_matrix_free_certificate_matches_dense_oracle(name, r, lam, tol, simple1_q, simple2_q):
    Q = {"simple1": simple1_q, "simple2": simple2_q, "nontight": nontight(), "syn": problems.synthetic_dense_q(120, seed=5)[0]}[name]
    R, s, primal = point(Q, r, lam, tol)
    ref = xo.certificate(Q, R * np.repeat(s, 3)[:, None], lam, primal)
    got = certificate.certify(lambda X: Q @ X, R, s, lam, primal)
    assert got["certified"] == ref["certified"]
    assert abs(got["min_eig"] - ref["min_eig"]) <= 1e-8 * max(1.0, got["lambda_max"])
    assert abs(got["dual"] - ref["dual"]) <= 1e-10 * max(1.0, abs(ref["dual"]))
    eig_tol = 1e-8 * max(1.0, got["lambda_max"])
    assert abs(got["gap"] - ref["gap"]) <= Q.shape[0] * eig_tol + 1e-9 * max(1.0, abs(ref["gap"]))   # gap carries 3N * min(0, lambda_min)
    assert got["matvecs"] < 5000
    if not ref["certified"]:         # the escape direction is the (simple) smallest eigenvector, up to sign
        assert min(np.abs(got["v"] - ref["v"]).max(), np.abs(got["v"] + ref["v"]).max()) < 1e-6


def test_closed_form_multipliers_solve_the_global_least_squares():
    rng = np.random.default_rng(2)
    N, r = 9, 4
    Q = nontight(N, 3)
    sR = rng.standard_normal((3 * N, r))
    Lam, y0 = certificate.multipliers(sR.reshape(N, 3, r), (Q @ sR).reshape(N, 3, r))
    A = xo._constraint_columns(sR)
    y, *_ = np.linalg.lstsq(A, (Q @ sR).reshape(-1, order="F"), rcond=None)
    np.testing.assert_allclose(y0, y[:6], atol=1e-10)
    fit_ref = A @ y
    fit = np.concatenate([Lam[i] @ sR[3 * i:3 * i + 3] for i in range(N)], axis=0).reshape(-1, order="F")
    np.testing.assert_allclose(fit, fit_ref, atol=1e-10)
    assert abs(np.trace(Lam[3])) < 1e-12 and np.allclose(Lam[3], Lam[3].T)          # cameras >= 1: symmetric traceless blocks


class OracleHandle:
    """Stand-in for capi.Handle backed by the oracle (CPU tests of the host logic only)."""

    def __init__(self, Q):
        self.Q = Q; self.N = Q.shape[0] // 3; self.is_bsr = False

    def comm_info(self):
        return {"rank": 0, "world": 1}

    def qy(self, X, alpha=1.0):
        return alpha * self.Q @ X

    def trust_region(self, R0, s0, lam=0.0, gradtol=1e-6, ls_step=0.0, v=None, max_time=1000.0):
        res = xo.trust_region(self.Q, xo.to_blocks(np.asarray(R0)), s0, lam, gradtol, ls_step, v, max_time)
        res.R = xo.from_blocks(res.Y)
        return res

    def certify(self, R, s, lam, primal):
        return xo.certificate(self.Q, R * np.repeat(s, 3)[:, None], lam, primal)


@pytest.mark.parametrize("method", ["dense", "lanczos"])
@pytest.mark.parametrize("name,max_rank,tol", [("simple1", 3, 1e-16), ("simple2", 5, 1e-1), ("nontight", 5, 1e-7), ("nontight", 4, 1e-7)])
def test_python_staircase_matches_oracle_staircase(name, max_rank, tol, method, simple1_q, simple2_q):
    Q = {"simple1": simple1_q, "simple2": simple2_q, "nontight": nontight()}[name]
    ref = xo.solve(Q, max_rank, tol, 0.0)
    out = solver.solve_arrays(OracleHandle(Q), max_rank, tol, 0.0, certificate_method=method)
    assert out["rank"] == ref["rank"] and out["status"] == ref["status"]
    assert len(out["trace"]) == len(ref["trace"])
    np.testing.assert_allclose(out["s"], ref["s"], atol=1e-6)
    assert abs(out["trace"][-1].primal - ref["trace"][-1].primal) <= 1e-7 * abs(ref["trace"][-1].primal)


def test_python_staircase_rank3_only_and_paths(tmp_path, simple1_q):
    from xm_code_b200 import binio
    binio.save_matrix_to_bin(str(tmp_path / "Q.bin"), simple1_q)
    h = OracleHandle(simple1_q)
    h.set_q_dense = lambda Q: None
    solver.solve_rank3(str(tmp_path), 3, 1e-6, 0.0, 1000, handle=h)
    R = binio.load_matrix_from_bin(str(tmp_path / "R.bin")); s = binio.load_matrix_from_bin(str(tmp_path / "s.bin"))
    ref = xo.solve(simple1_q, 3, 1e-6, 0.0, rank3_only=True)
    assert R.shape == (447, 3) and s.shape == (149, 1) and s[0, 0] == 1.0
    np.testing.assert_allclose(s[:, 0], ref["s"], atol=1e-9)
    assert solver.solve_rebuttle(str(tmp_path), 3, 1e-16, 0.0, 1000, handle=h) == 1


# ------------------------------------------------------------------------------------------------ CUDA path
@pytest.mark.gpu
def test_gpu_matrix_free_certificate_matches_dense(gpu_handle_factory, simple1_q):
    N = simple1_q.shape[0] // 3
    h = gpu_handle_factory()
    h.set_q_dense(simple1_q)
    got = h.trust_region(xo.from_blocks(xo.identity_init(N, 3)), np.ones(N), 0.0, 1e-16)
    dense = h.certify(got.R, got.s, 0.0, got.primal)
    free = certificate.certify(certificate.handle_operator(h), got.R, got.s, 0.0, got.primal)
    assert dense["certified"] and free["certified"]
    assert abs(free["min_eig"] - dense["min_eig"]) < 1e-7 and abs(free["dual"] - dense["dual"]) < 1e-8
    Q = nontight()
    h.set_q_dense(Q)
    r3 = h.trust_region(xo.from_blocks(xo.identity_init(30, 3)), np.ones(30), 0.0, 1e-8)
    dense = h.certify(r3.R, r3.s, 0.0, r3.primal)
    free = certificate.certify(certificate.handle_operator(h), r3.R, r3.s, 0.0, r3.primal)
    assert not dense["certified"] and not free["certified"]
    assert abs(free["min_eig"] - dense["min_eig"]) < 1e-8
    assert min(np.abs(free["v"] - dense["v"]).max(), np.abs(free["v"] + dense["v"]).max()) < 1e-6


@pytest.mark.gpu
@pytest.mark.parametrize("method", ["dense", "lanczos"])
def test_gpu_python_staircase_matches_oracle(gpu_handle_factory, method, simple2_q):
    for Q, max_rank, tol in ((simple2_q, 5, 1e-6), (nontight(), 5, 1e-7)):
        h = gpu_handle_factory()
        h.set_q_dense(Q)
        ref = xo.solve(Q, max_rank, tol, 0.0)
        out = solver.solve_arrays(h, max_rank, tol, 0.0, certificate_method=method)
        assert out["rank"] == ref["rank"] and out["status"] == ref["status"]
        np.testing.assert_allclose(out["s"], ref["s"], atol=1e-4)
        assert abs(out["trace"][-1].primal - ref["trace"][-1].primal) <= 1e-5 * abs(ref["trace"][-1].primal)
