"""CPU tests of the oracle itself: known answers on the reference's shipped inputs, derivative checks of the
restated operators, certificate, staircase.  (No GPU, no reference tree needed.)"""
import numpy as np
import pytest

from oracle import xm_oracle as xo


def rand_point(N, r, rng):
    Y = xo.mgs_rows(rng.standard_normal((N, 3, r)))
    s = np.concatenate([[1.0], rng.uniform(0.7, 1.4, N - 1)])
    return Y, s


def rand_psd(N, rng):
    A = rng.standard_normal((3 * N, 3 * N + 5))
    return A @ A.T / (3 * N)


def tangent(Y, s, rng):
    xi = rng.standard_normal(Y.shape)
    S = xo.sym3(np.einsum("iaj,ibj->iab", Y, xi))
    xi = xi - np.einsum("iab,ibj->iaj", S, Y)
    es = rng.standard_normal(s.shape); es[0] = 0.0
    return xi, es


def test_simple1_known_answer(simple1_q):
    # f* = 2.550991567720 (certified global optimum at rank 3), 13 outer iterations
    out = xo.solve(simple1_q, 3, 1e-16, 0.0)
    tr = out["trace"][0]
    assert out["status"] == 1 and out["rank"] == 3
    assert tr.outer_iters == 13
    assert abs(tr.primal - 2.550991567720) < 5e-11
    np.testing.assert_allclose(out["s"][1:6], [0.99764946, 1.0008994, 1.00081364, 1.00099953, 1.00061465], atol=2e-8)
    assert 0.99424 < out["s"].min() < 0.99425 and 1.00531 < out["s"].max() < 1.00532
    Y = xo.to_blocks(out["R"])
    assert np.allclose(np.linalg.det(Y), 1.0, atol=1e-9)
    np.testing.assert_allclose(np.einsum("iaj,ibj->iab", Y, Y), np.broadcast_to(np.eye(3), (Y.shape[0], 3, 3)), atol=1e-12)


def test_simple2_known_answer(simple2_q, simple2_obs):
    out = xo.solve(simple2_q, 5, 1e-10, 0.0)
    tr = out["trace"][-1]
    assert out["status"] == 1 and out["rank"] == 3
    assert abs(tr.primal - 4.8322430007e-02) < 1e-10
    # rotations against ground truth (pairwise relative rotation error, degrees)
    Y = xo.to_blocks(out["R"])
    gt = simple2_obs["gtR"]
    N = Y.shape[0]
    # the estimate is R_i^T stacked up to a global gauge; compare relative rotations R_i R_0^T
    est_rel = np.einsum("iaj,bj->iab", Y, Y[0])
    # ground truth file order differs from the solver order (most-observed frame first) -> only check orthogonality + det
    assert np.allclose(np.linalg.det(est_rel), 1.0, atol=1e-8)
    assert gt.shape == (3, 3 * N)


def test_gradient_matches_finite_difference():
    rng = np.random.default_rng(0)
    N, r, lam = 7, 4, 0.3
    Q = rand_psd(N, rng)
    Y, s = rand_point(N, r, rng)
    D, G, g = xo.egrad(Q, Y, s, lam)
    rgR, rgs = xo.project(Y, s, G, g)
    xi, es = tangent(Y, s, rng)
    f0 = xo.objective(Q, Y, s, lam)
    # metric: <a,b> = <aR,bR> + sum a_s b_s / s^2 ; rgs = s^2 g  =>  directional derivative = <rgR,xi> + g.es
    dd = float(np.vdot(rgR, xi)) + float(np.dot(rgs[1:] / s[1:] ** 2, es[1:]))
    for t in (1e-5, 1e-6):
        Yp, sp = xo.retract(Y, s, xi, es, t)
        Ym, sm = xo.retract(Y, s, xi, es, -t)
        fd = (xo.objective(Q, Yp, sp, lam) - xo.objective(Q, Ym, sm, lam)) / (2 * t)
        assert abs(fd - dd) < 1e-6 * max(1.0, abs(dd)), (fd, dd, f0)


def test_hessian_is_symmetric_in_the_metric():
    # <xi, Hess eta> == <eta, Hess xi> in the product metric — a strong check on ehess + ehess2rhess
    rng = np.random.default_rng(1)
    N, r, lam = 6, 5, 0.2
    Q = rand_psd(N, rng)
    Y, s = rand_point(N, r, rng)
    D, G, g = xo.egrad(Q, Y, s, lam)
    xi, xs = tangent(Y, s, rng)
    et, es = tangent(Y, s, rng)
    HxR, Hxs = xo.rhess_vec(Q, Y, s, lam, D, G, g, xi, xs)
    HeR, Hes = xo.rhess_vec(Q, Y, s, lam, D, G, g, et, es)
    a = float(np.vdot(et, HxR)) + float(np.dot(es[1:], Hxs[1:] / s[1:] ** 2))
    b = float(np.vdot(xi, HeR)) + float(np.dot(xs[1:], Hes[1:] / s[1:] ** 2))
    assert abs(a - b) < 1e-9 * max(1.0, abs(a))


def test_hessian_matches_second_difference():
    rng = np.random.default_rng(2)
    N, r, lam = 5, 3, 0.1
    Q = rand_psd(N, rng)
    Y, s = rand_point(N, r, rng)
    D, G, g = xo.egrad(Q, Y, s, lam)
    xi, xs = tangent(Y, s, rng)
    HR, Hs = xo.rhess_vec(Q, Y, s, lam, D, G, g, xi, xs)
    quad = float(np.vdot(xi, HR)) + float(np.dot(xs[1:], Hs[1:] / s[1:] ** 2))
    t = 1e-4

    def second_order_retract(tt):
        # the QR/MGS retraction is only first order; test the Hessian along a SECOND-order curve: polar factor
        # per camera (A A^T)^{-1/2} A for the Stiefel part, s*exp(eta/s) (the exact exponential) for the scales
        A = Y + tt * xi
        w, V = np.linalg.eigh(np.einsum("iaj,ibj->iab", A, A))
        inv_sqrt = np.einsum("iab,ib,icb->iac", V, 1.0 / np.sqrt(w), V)
        Yn = np.einsum("iab,ibj->iaj", inv_sqrt, A)
        sn = s.copy(); sn[1:] = s[1:] * np.exp(tt * xs[1:] / s[1:])
        return Yn, sn

    f0 = xo.objective(Q, Y, s, lam)
    fp = xo.objective(Q, *second_order_retract(t), lam)
    fm = xo.objective(Q, *second_order_retract(-t), lam)
    fd2 = (fp - 2 * f0 + fm) / t ** 2
    assert abs(fd2 - quad) < 1e-4 * max(1.0, abs(quad)), (fd2, quad)


def test_mgs_rows_is_orthonormal_with_positive_diagonal():
    rng = np.random.default_rng(3)
    A = rng.standard_normal((11, 3, 6))
    Qm = xo.mgs_rows(A)
    np.testing.assert_allclose(np.einsum("iaj,ibj->iab", Qm, Qm), np.broadcast_to(np.eye(3), (11, 3, 3)), atol=1e-13)
    assert np.all(np.einsum("iaj,iaj->ia", Qm, A) > 0)      # diag(R) > 0 : no sign fix needed (batchedQR.h)


def test_certificate_simple1(simple1_q):
    out = xo.solve(simple1_q, 3, 1e-16, 0.0)
    sR = out["R"] * np.repeat(out["s"], 3)[:, None]
    c = xo.certificate(simple1_q, sR, 0.0, out["trace"][0].primal)
    assert c["certified"] and c["min_eig"] > -1e-9 and abs(c["gap"]) < 1e-7


def test_rank_escalation_line_search_descends():
    # a point that is NOT second-order critical at rank 3: the line search along the escape direction must decrease f
    rng = np.random.default_rng(5)
    N = 12
    Q = rand_psd(N, rng)
    res3 = xo.trust_region(Q, xo.identity_init(N, 3), np.ones(N), 0.0, 1e-9)
    sR = xo.from_blocks(res3.Y * res3.s[:, None, None])
    c = xo.certificate(Q, sR, 0.0, res3.primal)
    Y0 = np.concatenate([res3.Y, np.zeros((N, 3, 1))], axis=2)
    v = (c["v"].reshape(N, 3) / res3.s[:, None]).reshape(-1)
    res4 = xo.trust_region(Q, Y0, res3.s, 0.0, 1e-9, ls_step=1.0, v=v)
    if not c["certified"]:
        assert res4.status == 0 and res4.primal <= res3.primal + 1e-12
    else:
        assert res4.status in (0, -1)


def test_bin_roundtrip(tmp_path):
    M = np.arange(12, dtype=np.float64).reshape(4, 3)
    xo.save_bin(tmp_path / "m.bin", M)
    np.testing.assert_array_equal(xo.load_bin(tmp_path / "m.bin"), M)


def test_e_recurrence_design_study_keeps_parity(simple1_q, simple2_q):
    """Design study for a tCG iteration with two barriers instead of three: updating the Euclidean Hessian
    product by the recurrence E <- beta E - 2 Q X(r_new) instead of recomputing 2 Q X(p_new) keeps the solver inside the
    parity bars (objective 1e-10, s 1e-8) and, while the problem is well conditioned, on the same trajectory."""
    from xm_code_b200 import problems
    for Q, r, lam, tol in ((simple1_q, 3, 0.0, 1e-16), (simple2_q, 3, 0.0, 1e-10), (problems.synthetic_dense_q(120, seed=5)[0], 5, 0.05, 1e-8)):
        N = Q.shape[0] // 3
        a = xo.trust_region(Q, xo.identity_init(N, r), np.ones(N), lam, tol)
        b = xo.trust_region(Q, xo.identity_init(N, r), np.ones(N), lam, tol, e_recurrence=True)
        assert abs(a.primal - b.primal) <= 1e-10 * abs(a.primal)
        np.testing.assert_allclose(a.s, b.s, atol=1e-8)
        assert a.outer_iters == b.outer_iters and abs(a.tcg_iters - b.tcg_iters) <= 0.02 * a.tcg_iters + 2
        n = len(a.log) - 3
        assert all(x[0] == y[0] and x[1] == y[1] and x[4] == y[4] and x[5] == y[5] for x, y in zip(a.log[:n], b.log[:n]))
