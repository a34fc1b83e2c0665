"""Certificate and rank staircase behind the C-ABI: xm_certify_ex (dense syevd / block Davidson on the Q.Y
operator), xm_op_diag_blocks, xm_solve — against the oracle's dense restatement of checkeig (global least squares + full
`eigh`, oracle/xm_oracle.py:certificate) and the oracle's own staircase (xo.solve).  GPU only: the product has no CPU path."""
import numpy as np
import pytest

from oracle import xm_oracle as xo
from xm_code_b200 import problems

pytestmark = pytest.mark.gpu


def nontight(N=30, seed=11):
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((3 * N, 3 * N + 2))
    return A @ A.T / (3 * N)


CASES = [("simple1", 3, 0.0, 1e-16), ("simple2", 3, 0.0, 1e-1), ("simple2", 3, 0.0, 1e-10), ("nontight", 3, 0.0, 1e-7),
         ("nontight200", 3, 0.0, 1e-7), ("syn", 3, 0.3, 1e-6), ("syn", 5, 0.05, 1e-7)]


@pytest.mark.parametrize("name,r,lam,tol", CASES)
def test_certificate_dense_and_iterative_match_the_oracle(name, r, lam, tol, simple1_q, simple2_q, gpu_handle_factory):
    Q = {"simple1": simple1_q, "simple2": simple2_q, "nontight": nontight(), "nontight200": nontight(200, 12),
         "syn": problems.synthetic_dense_q(120, seed=5)[0]}[name]
    N = Q.shape[0] // 3
    h = gpu_handle_factory()
    h.set_q_dense(Q)
    got = h.trust_region(xo.from_blocks(xo.identity_init(N, r)), np.ones(N), lam, tol)
    ref = xo.certificate(Q, got.R * np.repeat(got.s, 3)[:, None], lam, got.primal)
    lmax = float(np.linalg.eigvalsh(Q)[-1])
    for method in ("dense", "iterative"):
        c = h.certify(got.R, got.s, lam, got.primal, method=method)
        assert c["method"] == method
        assert c["certified"] == ref["certified"], (method, c, ref["min_eig"])
        eig_tol = 1e-8 * max(1.0, lmax)
        assert abs(c["min_eig"] - ref["min_eig"]) <= eig_tol, (method, c["min_eig"], ref["min_eig"])
        assert abs(c["dual"] - ref["dual"]) <= 1e-9 * max(1.0, abs(ref["dual"]))
        assert abs(c["gap"] - ref["gap"]) <= Q.shape[0] * eig_tol + 1e-9 * max(1.0, abs(ref["gap"]))     # gap carries 3N * min(0, lambda_min)
        if method == "iterative":
            assert c["converged"] and c["products"] <= 100, c          # <= 20 columns each: the bytes of <= 100 single products
        if not ref["certified"]:          # the escape direction is the (simple) smallest eigenvector, up to sign
            assert min(np.abs(c["v"] - ref["v"]).max(), np.abs(c["v"] + ref["v"]).max()) < 1e-5, method


def test_diag_blocks_dense_and_bsr(gpu_handle_factory):
    rng = np.random.default_rng(3)
    N = 157
    Q = rng.standard_normal((3 * N, 3 * N))
    h = gpu_handle_factory()
    h.set_q_dense(Q)
    D = h.diag_blocks()
    for i in (0, 1, 77, N - 1):
        np.testing.assert_array_equal(D[i], Q[3 * i:3 * i + 3, 3 * i:3 * i + 3])
    rowptr, col, vals = problems.erdos_renyi_bsr(200, avg_degree=7, seed=1)
    Qb = problems.bsr_to_dense(rowptr, col, vals)
    h.set_q_bsr(rowptr, col, vals, 3)
    D = h.diag_blocks()
    for i in (0, 5, 199):
        np.testing.assert_array_equal(D[i], Qb[3 * i:3 * i + 3, 3 * i:3 * i + 3])


@pytest.mark.parametrize("r", [3, 6])
def test_iterative_certificate_on_block_csr(gpu_handle_factory, r):
    """The dense route cannot run on a block-CSR operator; the iterative one gives the oracle's answer there (ER view graph)."""
    from xm_code_b200 import capi
    N = 260
    rowptr, col, vals = problems.erdos_renyi_bsr(N, avg_degree=8, seed=3)
    Q = problems.bsr_to_dense(rowptr, col, vals)
    h = gpu_handle_factory()
    h.set_q_bsr(rowptr, col, vals, 3)
    rng = np.random.default_rng(r)
    Y0 = xo.mgs_rows(rng.standard_normal((N, 3, r)))
    got = h.trust_region(xo.from_blocks(Y0), np.ones(N), 0.0, 1e-9)
    ref = xo.certificate(Q, got.R * np.repeat(got.s, 3)[:, None], 0.0, got.primal)
    c = h.certify(got.R, got.s, 0.0, got.primal)                      # auto -> iterative on block-CSR
    assert c["method"] == "iterative" and c["certified"] == ref["certified"] and c["converged"]
    assert abs(c["min_eig"] - ref["min_eig"]) <= 1e-8 * max(1.0, float(np.abs(Q).sum(axis=1).max()))
    with pytest.raises(capi.XmError):
        h.certify(got.R, got.s, 0.0, got.primal, method="dense")


@pytest.mark.parametrize("method", ["dense", "iterative"])
@pytest.mark.parametrize("name,max_rank,tol", [("simple1", 3, 1e-16), ("simple2", 5, 1e-6), ("nontight", 5, 1e-7), ("nontight", 4, 1e-7)])
def test_staircase_matches_oracle_staircase(name, max_rank, tol, method, simple1_q, simple2_q, gpu_handle_factory):
    """xm_solve (XM_main.cu:180-310): escalation, zero-padded column, v / s scaling, gradtol carry-over, statuses."""
    Q = {"simple1": simple1_q, "simple2": simple2_q, "nontight": nontight()}[name]
    ref = xo.solve(Q, max_rank, tol, 0.0)
    h = gpu_handle_factory()
    h.set_q_dense(Q)
    out = h.solve(max_rank, tol, 0.0, cert_method=method)
    assert out["rank"] == ref["rank"] and out["status"] == ref["status"]
    assert out["n_solves"] == len(ref["trace"]) and out["R"].shape == (Q.shape[0], ref["rank"])
    assert out["certificate_method"] == method
    np.testing.assert_allclose(out["s"], ref["s"], atol=1e-5)
    assert abs(out["primal"] - ref["trace"][-1].primal) <= 1e-6 * abs(ref["trace"][-1].primal)


def test_staircase_modes_and_path_api(tmp_path, simple1_q, gpu_handle_factory):
    from xm_code_b200 import binio, solver
    binio.save_matrix_to_bin(str(tmp_path / "Q.bin"), simple1_q)
    h = gpu_handle_factory()
    solver.solve_rank3(str(tmp_path), 3, 1e-6, 0.0, 1000, handle=h)
    R = binio.load_matrix_from_bin(str(tmp_path / "R.bin")); s = binio.load_matrix_from_bin(str(tmp_path / "s.bin"))
    ref = xo.solve(simple1_q, 3, 1e-6, 0.0, rank3_only=True)
    assert R.shape == (447, 3) and s.shape == (149, 1) and s[0, 0] == 1.0
    np.testing.assert_allclose(s[:, 0], ref["s"], atol=1e-7)
    assert solver.solve_rebuttle(str(tmp_path), 3, 1e-16, 0.0, 1000, handle=h) == 1
    binio.save_matrix_to_bin(str(tmp_path / "s_ini.bin"), np.full((149, 1), 1.01))
    assert solver.solve_rebuttle(str(tmp_path), 3, 1e-12, 0.0, 1000, handle=h) == 1       # s_ini is used, R_ini is not (XM_main.cu:95-103)
    solver.solve(str(tmp_path), 3, 1e-16, 0.0, 1000, handle=h)
    s2 = binio.load_matrix_from_bin(str(tmp_path / "s.bin"))
    np.testing.assert_allclose(s2[:, 0], xo.solve(simple1_q, 3, 1e-16, 0.0)["s"], atol=1e-8)
