"""The compiled (C + OpenMP) oracle oracle/xm_oracle_c.c: pinned against OUTPUT OF THE REFERENCE ITSELF (tests/golden/ref_*,
same checks as tests/test_oracle_vs_reference.py applies to the NumPy oracle), cross-checked against the NumPy oracle on
seeded problems, and independent of the thread count.  CPU only."""
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import xm_oracle as xo, xm_oracle_c as xc
from conftest import GOLD, ROOT, anchored_gram, load_bin
from test_oracle_vs_reference import case_dir, compare_trace, parse_reference_log


@pytest.mark.parametrize("name,tol,q_source", [("simple1", 1e-16, "simple1"), ("simple2", 1e-10, "simple2"), ("syn100", 1e-6, "syn100")])
def test_c_oracle_rank3_solves_match_the_reference(name, tol, q_source, simple1_q, simple2_q):
    d = case_dir(name)
    if q_source == "simple1":
        Q = simple1_q
    elif q_source == "simple2":
        Q = simple2_q
    else:
        from xm_code_b200 import problems
        Q, _ = problems.synthetic_dense_q(100, seed=1)
    rows, total, js = parse_reference_log(os.path.join(d, "log.txt"))
    N = Q.shape[0] // 3
    res = xc.trust_region(Q, xo.identity_init(N, 3), np.ones(N), 0.0, tol)
    # identical iteration structure except for the last, rounding-dominated iterations (the summation order of the Q.Y
    # rows differs from cuBLAS / OpenBLAS, so the C twin leaves the common trajectory one outer iteration earlier than NumPy)
    compare_trace(rows, res.log, n_exact=max(1, len(rows) - 4))
    assert abs(total - res.tcg_iters) <= 0.05 * total
    primal_ref = js["runs"][-1]["primal"]
    assert abs(res.primal - primal_ref) <= 1e-10 * abs(primal_ref)
    R_ref = load_bin(os.path.join(d, "R_ref.bin")); s_ref = load_bin(os.path.join(d, "s_ref.bin"))[:, 0]
    np.testing.assert_allclose(res.s, s_ref, atol=1e-8 if tol <= 1e-9 else 1e-5)
    np.testing.assert_allclose(anchored_gram(xo.from_blocks(res.Y), res.s), anchored_gram(R_ref, s_ref), atol=1e-7 if tol <= 1e-9 else 1e-4)


def test_c_oracle_rank_escalation_matches_the_reference():
    d = case_dir("esc30_r4")
    rng = np.random.default_rng(11); N = 30
    A = rng.standard_normal((3 * N, 3 * N + 2)); Q = A @ A.T / (3 * N)
    R0 = load_bin(os.path.join(d, "R_ini.bin")); s0 = load_bin(os.path.join(d, "s_ini.bin"))[:, 0]; v = load_bin(os.path.join(d, "v_ini.bin"))[:, 0]
    rows, total, js = parse_reference_log(os.path.join(d, "log.txt"))
    res = xc.trust_region(Q, xo.to_blocks(R0), s0, 0.0, 1e-7, ls_step=1.0, v=v)
    assert res.status == 0
    compare_trace(rows[:6], res.log[:6], n_exact=6)
    primal_ref = js["runs"][-1]["primal"]
    assert abs(res.primal - primal_ref) <= 1e-7 * abs(primal_ref)
    bad = xc.trust_region(Q, xo.to_blocks(R0), s0, 0.0, 1e-7, ls_step=1.0, v=np.full(3 * N, np.nan))
    assert bad.status == -1 and bad.primal == -1.0                         # line-search failure exit (:384-405)


@pytest.mark.parametrize("N,r,lam,tol", [(40, 3, 0.0, 1e-8), (80, 3, 0.05, 1e-8), (60, 5, 0.2, 1e-7), (700, 3, 0.0, 1e-6)])
def test_c_oracle_agrees_with_numpy_oracle(N, r, lam, tol):
    from xm_code_b200 import problems
    Q, _ = problems.synthetic_dense_q(N, seed=N + r)
    Y0 = xo.identity_init(N, r)
    a = xo.trust_region(Q, Y0, np.ones(N), lam, tol)
    b = xc.trust_region(Q, Y0, np.ones(N), lam, tol)
    assert abs(a.primal - b.primal) <= 1e-9 * abs(a.primal)
    np.testing.assert_allclose(a.s, b.s, atol=1e-6)
    assert b.gradtol in (pytest.approx(tol), pytest.approx(tol / 10))     # quirk Q1; which exit fires at the end is rounding-dependent
    n = min(len(a.log), len(b.log), 6)                                    # identical early trajectory
    for x, y in zip(a.log[:n], b.log[:n]):
        assert x[0] == y[0] and x[1] == y[1] and x[4] == y[4] and x[5] == y[5]
        assert abs(x[2] - y[2]) <= 1e-10 * abs(x[2])


def test_c_oracle_is_independent_of_the_thread_count(tmp_path):
    from xm_code_b200 import problems
    Q, _ = problems.synthetic_dense_q(650, seed=3)           # built ONCE: OpenBLAS's own rounding depends on OMP_NUM_THREADS
    np.save(tmp_path / "Q.npy", Q)
    code = ("import sys; sys.path.insert(0, %r); import numpy as np\n"
            "from oracle import xm_oracle as xo, xm_oracle_c as xc\n"
            "Q = np.load(%r)\n"
            "r = xc.trust_region(Q, xo.identity_init(650, 3), np.ones(650), 0.0, 1e-6)\n"
            "print(xc.num_threads(), repr(r.primal), r.tcg_iters, repr(float(r.s.sum())))\n") % (ROOT, str(tmp_path / "Q.npy"))
    outs = []
    for nt in ("1", "4"):
        env = dict(os.environ, OMP_NUM_THREADS=nt)
        out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=600)
        assert out.returncode == 0, out.stderr[-1500:]
        outs.append(out.stdout.split())
    assert outs[0][0] == "1" and outs[1][0] == "4"
    assert outs[0][1:] == outs[1][1:]                                      # bit-identical: reductions are serial per-camera sums
