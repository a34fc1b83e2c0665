"""GPU parity cases added in round 2 (VERDICT r1 "close the parity gaps"), all through the C-ABI:
  * the headline-sized problem C3 (BAL-Ladybug-1723-shaped) against OUTPUT OF THE REFERENCE ITSELF
    (tests/golden/ref_bal1723: its stdout table, final objective and scales) and against the compiled C oracle;
  * single-GPU block-CSR SOLVES (xm_solve_kernel<.,512,2>) at ranks 4, 5, 10 and 20 against the oracle;
  * dense full solves at ranks 12 and 20 (the 256-thread instantiations)."""
import os

import numpy as np
import pytest

from oracle import xm_oracle as xo
from conftest import GOLD, anchored_gram, load_bin
from test_oracle_vs_reference import parse_reference_log

pytestmark = pytest.mark.gpu


def rand_point(N, r, rng):
    Y = xo.mgs_rows(rng.standard_normal((N, 3, r)))
    s = np.concatenate([[1.0], rng.uniform(0.8, 1.25, N - 1)])
    return Y, s


def check_point(got, ref, primal_rel, s_abs, x_abs):
    assert abs(got.primal - ref.primal) <= primal_rel * abs(ref.primal)
    np.testing.assert_allclose(got.s, ref.s, atol=s_abs, rtol=0)
    np.testing.assert_allclose(anchored_gram(got.R, got.s), anchored_gram(xo.from_blocks(ref.Y), ref.s), atol=x_abs, rtol=0)


def test_c3_bal1723_matches_reference_output_and_c_oracle(gpu_handle_factory):
    """BASELINE config 3 at full size.  Golden: the unmodified reference trustregion.h run on a B200 on this very Q
    (oracle/make_ref_goldens.sh); tolerance: objective 1e-10 relative, scales 1e-6 (the solve stops at gradnorm < 1e-6),
    iteration table identical while the problem is far from convergence, total tCG iterations within 5 %."""
    from xm_code_b200 import problems
    from oracle import xm_oracle_c as xc
    N = 1723
    Q, _ = problems.synthetic_dense_q(N, seed=0, obs_per_camera=60, n_landmarks=12 * N)
    rows, total, js = parse_reference_log(os.path.join(GOLD, "ref_bal1723", "log.txt"))
    restart = [n for n, row in enumerate(rows) if row[0] == 0]           # the golden log holds three identical repeats of the solve
    rows = rows[:restart[1]] if len(restart) > 1 else rows
    s_ref = load_bin(os.path.join(GOLD, "ref_bal1723", "s_ref.bin"))[:, 0]
    primal_ref = js["runs"][-1]["primal"]
    # the reference's order of operations inside a tCG iteration (three barriers): iteration counts within 5 % of the reference run;
    # the default two-barrier iteration (E recurrence) reaches the same optimum on a slightly different rounding path
    h3 = gpu_handle_factory(three_barrier_tcg=True)
    h3.set_q_dense(Q)
    got3 = h3.trust_region(xo.from_blocks(xo.identity_init(N, 3)), np.ones(N), 0.0, 1e-6)
    assert abs(got3.primal - primal_ref) <= 1e-10 * abs(primal_ref) and abs(got3.stats["tcg_iters"] - total) <= 0.05 * total
    np.testing.assert_allclose(got3.s, s_ref, atol=1e-6, rtol=0)
    h = gpu_handle_factory()
    h.set_q_dense(Q)
    got = h.trust_region(xo.from_blocks(xo.identity_init(N, 3)), np.ones(N), 0.0, 1e-6)
    # the last outer iterations are rounding-dominated: the solve ends by the small-gradient test (like the reference run) or one
    # superlinear tCG solve earlier by the rdotr < 1e-15 test; both are KKT points to 1e-6
    assert got.stats["exit"] in ("gradtol", "rdotr_tiny") and got.stats["gradnorm"] < 1e-6
    assert abs(got.primal - primal_ref) <= 1e-10 * abs(primal_ref)
    np.testing.assert_allclose(got.s, s_ref, atol=1e-6, rtol=0)
    # the reference's own table: same structure for the first rows (rounding decides accept/reject ties later on)
    n_exact = 12
    for a, b in zip(rows[:n_exact], got.log[:n_exact]):
        assert a[0] == b[0] and a[1] == b[1] and a[4] == b[4] and a[5] == b[5], (a, b)
        assert abs(a[2] - b[2]) <= 6e-4 * abs(b[2]) and abs(a[3] - b[3]) <= 6e-4 * abs(b[3]), (a, b)
    assert abs(got.stats["tcg_iters"] - total) <= 0.15 * total
    assert abs(len(got.log) - len(rows)) <= 0.20 * len(rows)
    # and the compiled oracle (bit-for-bit the NumPy oracle's arithmetic, ~4 s on 16 cores)
    ref = xc.trust_region(np.ascontiguousarray(Q), xo.identity_init(N, 3), np.ones(N), 0.0, 1e-6)
    check_point(got, ref, primal_rel=1e-10, s_abs=1e-6, x_abs=1e-5)
    # iteration counts are decided by rounding after ~1000 iterations: reference 1474, C oracle 1406, this kernel 1525 / 1356
    assert abs(got3.stats["tcg_iters"] - ref.tcg_iters) <= 0.12 * ref.tcg_iters


@pytest.mark.parametrize("r", [4, 5, 10, 20])
def test_bsr_solve_single_gpu_matches_oracle(gpu_handle_factory, r):
    """The persistent block-CSR solve kernel on ONE GPU (Erdos-Renyi view graph, 320 cameras) from a random rank-r point."""
    from xm_code_b200 import problems
    N = 320
    rowptr, col, vals = problems.erdos_renyi_bsr(N, avg_degree=9, seed=6)
    Q = problems.bsr_to_dense(rowptr, col, vals)
    rng = np.random.default_rng(40 + r)
    Y0, s0 = rand_point(N, r, rng)
    ref = xo.trust_region(Q, Y0, s0, 0.0, 1e-8)
    h = gpu_handle_factory()
    h.set_q_bsr(rowptr, col, vals, 3)
    got = h.trust_region(xo.from_blocks(Y0), s0, 0.0, 1e-8)
    assert got.stats["exit"] in ("gradtol", "rdotr_tiny")
    check_point(got, ref, primal_rel=1e-8, s_abs=1e-6, x_abs=1e-5)
    # first outer iterations: identical structure and numbers (same arithmetic up to summation order)
    for a, b in zip(got.log[:4], ref.log[:4]):
        assert a[0] == b[0] and a[1] == b[1]
        assert abs(a[2] - b[2]) <= 1e-10 * abs(b[2]) and abs(a[3] - b[3]) <= 1e-8 * abs(b[3])
    # same operator through the dense kernel: both CUDA paths agree with each other as well
    hd = gpu_handle_factory()
    hd.set_q_dense(Q)
    gd = hd.trust_region(xo.from_blocks(Y0), s0, 0.0, 1e-8)
    assert abs(gd.primal - got.primal) <= 1e-8 * abs(got.primal)


@pytest.mark.parametrize("r,lam", [(7, 0.0), (10, 0.05), (12, 0.0), (20, 0.05)])
def test_dense_solve_high_rank_matches_oracle(gpu_handle_factory, r, lam):
    """Dense full solves on the 256-thread instantiations from a random rank-r point: padded ranks 8 / 10 (two cameras per consumer
    warp: 151 cameras on 148 CTAs and odd batch sizes exercise the half-filled warp) and 12 / 20 (one camera per warp)."""
    from xm_code_b200 import problems
    N = 151
    Q, _ = problems.synthetic_dense_q(N, seed=7)
    rng = np.random.default_rng(50 + r)
    Y0, s0 = rand_point(N, r, rng)
    ref = xo.trust_region(Q, Y0, s0, lam, 1e-8)
    h = gpu_handle_factory()
    h.set_q_dense(Q)
    got = h.trust_region(xo.from_blocks(Y0), s0, lam, 1e-8)
    assert got.stats["threads_per_cta"] == 256
    check_point(got, ref, primal_rel=1e-8, s_abs=1e-6, x_abs=1e-5)
    for a, b in zip(got.log[:4], ref.log[:4]):
        assert a[0] == b[0] and a[1] == b[1]
        assert abs(a[2] - b[2]) <= 1e-10 * abs(b[2]) and abs(a[3] - b[3]) <= 1e-8 * abs(b[3])
