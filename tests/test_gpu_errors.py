"""GPU: error behaviour of the C-ABI (include/xm_b200.h).  The reference prints CUDA errors and carries on
(XM/include/Utils/check.h:41-76); the C-ABI returns a negative code and a message instead, and the
algorithm's own numerical exits are reported in xm_stats.exit_code, not as errors."""
import ctypes as C

import numpy as np
import pytest

from oracle import xm_oracle as xo
from xm_code_b200 import capi

pytestmark = pytest.mark.gpu


def test_calls_before_a_matrix_is_set_are_refused(gpu_handle_factory):
    h = gpu_handle_factory()
    with pytest.raises(capi.XmError, match="XM_EINVAL"):
        h.qy(np.zeros((6, 3)))
    with pytest.raises(capi.XmError, match="XM_EINVAL"):
        h.trust_region(np.zeros((6, 3)), np.ones(2))


def test_bad_shapes_and_ranks_are_refused(gpu_handle_factory):
    h = gpu_handle_factory()
    lib = h.lib
    Q = np.eye(6)
    assert lib.xm_set_q_dense(h._h, 5, Q.ctypes.data, 5) == -1            # 3N must be a multiple of 3
    assert lib.xm_set_q_dense(h._h, 6, Q.ctypes.data, 4) == -1            # leading dimension shorter than the matrix
    assert lib.xm_set_q_dense(h._h, 6, None, 6) == -1
    h.set_q_dense(Q)
    for r in (1, 2, 21):                                                   # rank outside [3, XM_MAX_RANK]
        with pytest.raises(capi.XmError, match="XM_EINVAL"):
            h.qy(np.zeros((6, r)))
    gt = C.c_double(1e-6)
    R0 = np.asfortranarray(xo.from_blocks(xo.identity_init(2, 3))); s0 = np.ones(2)
    # a line-search call (ls_step != 0) without a direction is refused rather than dereferenced
    rc = lib.xm_trust_region(h._h, 3, R0.ctypes.data, s0.ctypes.data, 0.0, C.byref(gt), 1.0, None, 10.0, R0.ctypes.data, s0.ctypes.data, None, None, None)
    assert rc == -1 and b"null" in lib.xm_last_error(h._h)
    assert lib.xm_set_q_bsr(h._h, 2, 5, None, None, None) == -1           # block size must be 3 or 4
    assert lib.xm_comm_init(h._h, 0, 9, 100, 3, None) == -1               # more ranks than XM_MAX_WORLD
    assert lib.xm_comm_init(h._h, 2, 2, 100, 3, None) == -1               # rank outside the world


def test_numerical_exits_are_reported_not_raised(gpu_handle_factory):
    """A direction whose trial objective is never below f0 ends the rank-escalation line search: primal = -1 and
    exit_code = XM_EXIT_LINESEARCH_FAILED (trustregion.h:384-405), return code XM_OK."""
    rng = np.random.default_rng(3)
    N = 12
    A = rng.standard_normal((3 * N, 3 * N + 2)); Q = A @ A.T / (3 * N)
    h = gpu_handle_factory()
    h.set_q_dense(Q)
    Y0 = np.concatenate([xo.identity_init(N, 3), np.zeros((N, 3, 1))], axis=2)
    res = h.trust_region(xo.from_blocks(Y0), np.ones(N), 0.0, 1e-6, ls_step=1.0, v=np.full(3 * N, np.nan))
    assert res.primal == -1.0 and res.stats["exit"] == "linesearch_failed"
    # the iteration caps are honoured and reported
    h2 = gpu_handle_factory(max_outer=2)
    h2.set_q_dense(Q)
    res = h2.trust_region(xo.from_blocks(xo.identity_init(N, 3)), np.ones(N), 0.0, 1e-12)
    assert res.stats["exit"] == "max_outer" and res.stats["outer_iters"] == 2


def test_tiny_sizes(gpu_handle_factory):
    """Two and three cameras: the smallest problems with a scale degree of freedom."""
    rng = np.random.default_rng(5)
    for N in (2, 3):
        A = rng.standard_normal((3 * N, 3 * N + 1)); Q = A @ A.T
        h = gpu_handle_factory()
        h.set_q_dense(Q)
        got = h.trust_region(xo.from_blocks(xo.identity_init(N, 3)), np.ones(N), 0.0, 1e-10)
        ref = xo.trust_region(Q, xo.identity_init(N, 3), np.ones(N), 0.0, 1e-10)
        assert got.s[0] == 1.0
        assert abs(got.primal - ref.primal) <= 1e-9 * abs(ref.primal)
        np.testing.assert_allclose(got.s, ref.s, atol=1e-7)
