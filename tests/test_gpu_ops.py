"""GPU parity, op level: every device phase against the oracle on seeded random inputs, through the C-ABI.
Tolerance for this FP64 path: 1e-12 relative; the operation order differs from the reference's
cuBLAS calls, so bit-exactness is not defined."""
import numpy as np
import pytest

from oracle import xm_oracle as xo

pytestmark = pytest.mark.gpu
TOL = 1e-12


def rel(a, b):
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b))) / max(1e-300, np.max(np.abs(b))))


def rand_psd(n3, rng):
    A = rng.standard_normal((n3, n3 + 3))
    return A @ A.T / n3


def rand_point(N, r, rng):
    Y = xo.mgs_rows(rng.standard_normal((N, 3, r)))
    s = np.concatenate([[1.0], rng.uniform(0.7, 1.4, N - 1)])
    return Y, s


CASES = [(1, 3), (2, 3), (5, 4), (33, 3), (64, 5), (149, 3), (150, 7), (301, 10), (77, 13), (40, 20), (500, 3), (700, 6)]


@pytest.mark.parametrize("N,r", CASES)
def test_qy_matches_oracle(gpu_handle_factory, N, r):
    rng = np.random.default_rng(100 + N + r)
    Q = rand_psd(3 * N, rng)
    Q = Q + 0.1 * rng.standard_normal(Q.shape)        # deliberately NOT symmetric: out = Q X, not Q^T X
    h = gpu_handle_factory()
    h.set_q_dense(Q)
    X = rng.standard_normal((3 * N, r))
    got = h.qy(X, alpha=2.0)
    assert rel(got, 2.0 * Q @ X) < TOL


@pytest.mark.parametrize("variant", [0, 1])
@pytest.mark.parametrize("grid,ks", [(1, 1), (1, 16), (3, 4), (16, 2), (148, 0), (40, 8), (97, 15), (20, 3)])
def test_qy_launch_geometries(gpu_handle_factory, grid, ks, variant):
    """Both dense Q.Y paths (0: 2-D TMA ring, 1: direct streaming loads) over grid sizes / k-splits, incl. multi-batch
    CTAs (grid=1: 97 cameras in one CTA), ring wrap-around and the OOB-filled tail chunk (3N = 291 is not a multiple of 128)."""
    rng = np.random.default_rng(7)
    N, r = 97, 5
    Q = rand_psd(3 * N, rng)
    X = rng.standard_normal((3 * N, r))
    h = gpu_handle_factory(grid_ctas=grid, ksplit=ks, qy_variant=variant)
    h.set_q_dense(Q)
    assert rel(h.qy(X), Q @ X) < TOL
    assert rel(h.qy(2 * X), 2 * Q @ X) < TOL          # second call on the same handle (ring state is per launch)


@pytest.mark.parametrize("r", [8, 10])
@pytest.mark.parametrize("grid,ks", [(1, 1), (1, 7), (3, 0), (16, 0), (148, 0), (40, 7), (97, 0)])
def test_qy_two_cameras_per_warp_geometries(gpu_handle_factory, grid, ks, r):
    """Padded ranks 8 / 10 run 256-thread CTAs whose consumer warps sweep TWO cameras per operand load: grid sizes / k-splits with
    odd batch sizes (a warp's second camera missing), multi-batch CTAs and a single camera per CTA."""
    rng = np.random.default_rng(17)
    N = 97
    Q = rand_psd(3 * N, rng) + 0.1 * rng.standard_normal((3 * N, 3 * N))
    X = rng.standard_normal((3 * N, r))
    h = gpu_handle_factory(grid_ctas=grid, ksplit=ks)
    h.set_q_dense(Q)
    assert rel(h.qy(X), Q @ X) < TOL
    assert rel(h.qy(-X, alpha=0.5), -0.5 * Q @ X) < TOL


@pytest.mark.parametrize("variant", [0, 1])
def test_solver_paths_agree(gpu_handle_factory, variant):
    from xm_code_b200 import problems
    Q, _ = problems.synthetic_dense_q(200, seed=9)
    ref = xo.trust_region(Q, xo.identity_init(200, 3), np.ones(200), 0.0, 1e-7)
    for grid in (0, 5):                                  # 5 CTAs x 40 cameras: multi-batch, prefetch across batches
        h = gpu_handle_factory(qy_variant=variant, grid_ctas=grid)
        h.set_q_dense(Q)
        got = h.trust_region(xo.from_blocks(xo.identity_init(200, 3)), np.ones(200), 0.0, 1e-7)
        assert abs(got.primal - ref.primal) <= 1e-9 * abs(ref.primal)
        assert np.max(np.abs(got.s - ref.s)) < 1e-6


def test_qy_strided_ld_and_reupload(gpu_handle_factory):
    rng = np.random.default_rng(8)
    N = 20
    big = rng.standard_normal((3 * N + 5, 3 * N))
    bigF = np.asfortranarray(big)                      # keep alive: the pointer below refers to it
    Q = bigF[: 3 * N, :]                                # column-major view with ld = 3N + 5
    h = gpu_handle_factory()
    h.set_q_dense_ptr(3 * N, bigF.ctypes.data, ld=3 * N + 5)
    X = rng.standard_normal((3 * N, 3))
    assert rel(h.qy(X), Q @ X) < TOL
    Q2 = rand_psd(3 * 31, rng)                          # different size on the same handle
    h.set_q_dense(Q2)
    X2 = rng.standard_normal((93, 4))
    assert rel(h.qy(X2), Q2 @ X2) < TOL


@pytest.mark.parametrize("N,r,lam", [(2, 3, 0.0), (9, 3, 0.5), (64, 5, 0.0), (149, 4, 0.1), (333, 8, 0.0), (50, 16, 0.2)])
def test_objective_gradient_hessian_retraction(gpu_handle_factory, N, r, lam):
    rng = np.random.default_rng(200 + N + r)
    Q = rand_psd(3 * N, rng)
    Y, s = rand_point(N, r, rng)
    R = xo.from_blocks(Y)
    h = gpu_handle_factory()
    h.set_q_dense(Q)
    # objective (trustregion.h:162-170)
    f = h.objective(R, s, lam)
    assert abs(f - xo.objective(Q, Y, s, lam)) < TOL * abs(f) * 10
    # Riemannian gradient (grad :186-194 + projection :307-317)
    D, G, g = xo.egrad(Q, Y, s, lam)
    rgR, rgs = xo.project(Y, s, G, g)
    gR, gs, gn = h.rgrad(R, s, lam)
    assert rel(gR, xo.from_blocks(rgR)) < TOL * 10
    assert rel(gs, rgs) < TOL * 10 and gs[0] == 0.0
    ref_gn = np.sqrt(np.vdot(rgR, rgR) + np.sum((rgs[1:] / s[1:]) ** 2))
    assert abs(gn - ref_gn) < TOL * 10 * ref_gn
    # Riemannian Hessian-vector (ehess :227-255 + ehess2rhess :277-295)
    P = rng.standard_normal(Y.shape)
    ps = rng.standard_normal(N); ps[0] = 0.0
    HR, Hs = xo.rhess_vec(Q, Y, s, lam, D, G, g, P, ps)
    gHR, gHs = h.rhess(R, s, xo.from_blocks(P), ps, lam)
    assert rel(gHR, xo.from_blocks(HR)) < TOL * 10
    assert rel(gHs, Hs) < TOL * 10
    # retraction (:341-351, batchedQR.h:42-67)
    eta = 0.3 * rng.standard_normal(Y.shape); es = 0.2 * rng.standard_normal(N)
    Yn, sn = xo.retract(Y, s, eta, es, 0.7)
    gRn, gsn = h.retract(R, s, xo.from_blocks(eta), es, 0.7)
    assert rel(gRn, xo.from_blocks(Yn)) < TOL * 10
    assert rel(gsn, sn) < TOL * 10 and gsn[0] == 1.0


def test_bsr_operator_matches_dense(gpu_handle_factory):
    from xm_code_b200 import problems
    rowptr, col, vals = problems.erdos_renyi_bsr(120, avg_degree=10, seed=4)
    Q = problems.bsr_to_dense(rowptr, col, vals)
    rng = np.random.default_rng(3)
    h = gpu_handle_factory()
    h.set_q_bsr(rowptr, col, vals, 3)
    for r in (3, 5, 10):
        X = rng.standard_normal((360, r))
        assert rel(h.qy(X, 0.5), 0.5 * Q @ X) < TOL
    Y, s = rand_point(120, 4, rng)
    assert abs(h.objective(xo.from_blocks(Y), s, 0.0) - xo.objective(Q, Y, s, 0.0)) < 1e-11 * abs(xo.objective(Q, Y, s, 0.0))


@pytest.mark.parametrize("r", [3, 5, 8, 20])
def test_bsr_ragged_rows_and_4x4_blocks(gpu_handle_factory, r):
    """Block-CSR edge cases through both thread geometries (512 threads at r <= 5, 1024 above): empty block rows, single-block rows,
    row lengths at and around the staged chunk sizes (16 / 32 blocks) and far beyond them, unsorted column order inside a row, and the
    4x4-block input form (one 4x4 block per view-graph edge: only the leading 3x3 acts on the rotation rows)."""
    rng = np.random.default_rng(60 + r)
    N = 230
    lengths = [0, 1, 15, 16, 17, 31, 32, 33, 47, 48, 49, 64, 65, 2, 130, 0, 229]
    rowptr = [0]; col = []
    for i in range(N):
        k = lengths[i % len(lengths)]
        col.extend(rng.permutation(N)[:k].tolist())            # deliberately unsorted
        rowptr.append(len(col))
    rowptr = np.array(rowptr, dtype=np.int32); col = np.array(col, dtype=np.int32)
    blocks = rng.standard_normal((col.size, 3, 3))               # blocks[b][row][col]
    Q = np.zeros((3 * N, 3 * N))
    for i in range(N):
        for b in range(rowptr[i], rowptr[i + 1]):
            Q[3 * i:3 * i + 3, 3 * col[b]:3 * col[b] + 3] += blocks[b]          # (a permutation has no duplicate columns)
    X = rng.standard_normal((3 * N, r))
    h = gpu_handle_factory()
    h.set_q_bsr(rowptr, col, np.ascontiguousarray(np.swapaxes(blocks, 1, 2)), 3)     # wire format: column-major inside a block
    got = h.qy(X, alpha=1.5)
    assert rel(got, 1.5 * Q @ X) < TOL
    empty = [i for i in range(N) if rowptr[i] == rowptr[i + 1]]
    assert empty
    for i in empty:
        assert np.all(got[3 * i:3 * i + 3] == 0.0)
    b4 = np.zeros((col.size, 4, 4)); b4[:, :3, :3] = blocks
    b4[:, 3, :] = rng.standard_normal((col.size, 4)); b4[:, :, 3] = rng.standard_normal((col.size, 4))     # translation parts: ignored
    h4 = gpu_handle_factory()
    h4.set_q_bsr(rowptr, col, np.ascontiguousarray(np.swapaxes(b4, 1, 2)), 4)
    assert rel(h4.qy(X, alpha=1.5), 1.5 * Q @ X) < TOL
