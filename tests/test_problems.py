"""Host-side logic on CPU: Q assembly vs the reference's create_matrix golden, .bin I/O, BSR generator."""
import numpy as np

from xm_code_b200 import binio, problems


def test_q_from_observations_matches_reference_create_matrix(simple2_obs, simple2_q):
    o = simple2_obs
    Q = problems.q_from_observations(int(o["N"]), int(o["M"]), o["edges"][:, 0] - 1, o["edges"][:, 1] - 1, o["weights"], o["pts"])
    assert Q.shape == simple2_q.shape == (279, 279)
    assert np.abs(Q - simple2_q).max() / np.abs(simple2_q).max() < 1e-10


def test_synthetic_q_is_psd_with_gauge_nullspace():
    Q, prob = problems.synthetic_dense_q(40, seed=2, noise=0.0)
    w = np.linalg.eigvalsh(Q)
    assert w[0] > -1e-8 * w[-1]
    # noise-free: the ground truth U = [s_i R_i] has zero cost  ->  rows of U^T span the null space
    U = np.concatenate([prob["s"][i] * prob["R"][i].T for i in range(prob["N"])], axis=0)     # 3N x 3
    assert np.abs(Q @ U).max() < 1e-8 * w[-1]


def test_binio_roundtrip(tmp_path):
    M = np.random.default_rng(0).standard_normal((6, 4))
    binio.save_matrix_to_bin(str(tmp_path / "a.bin"), M)
    np.testing.assert_array_equal(binio.load_matrix_from_bin(str(tmp_path / "a.bin")), M)
    raw = open(tmp_path / "a.bin", "rb").read()
    assert int.from_bytes(raw[:4], "little") == 6 and int.from_bytes(raw[4:8], "little") == 4
    np.testing.assert_array_equal(np.frombuffer(raw[8:], dtype=np.float64), M.ravel(order="F"))


def test_erdos_renyi_bsr_is_symmetric_psd():
    rowptr, col, vals = problems.erdos_renyi_bsr(30, avg_degree=6, seed=1)
    Q = problems.bsr_to_dense(rowptr, col, vals)
    assert np.abs(Q - Q.T).max() < 1e-12
    assert np.linalg.eigvalsh(Q)[0] > 0


def test_host_assembly_matches_reference_q_and_abar():
    """The host restatement of the assembly (problems.q_from_observations: what the generators and the GPU tests of
    xm_create_matrix compare against) against Q.bin / Abar.bin written by the reference's own utils/creatematrix.create_matrix for
    the same observations (tests/golden/recover_ref.npz).  The drop-in create_matrix itself runs on the GPU: tests/test_gpu_assemble.py."""
    import os
    from conftest import GOLD
    g = np.load(os.path.join(GOLD, "recover_ref.npz"))
    prob = problems.synthetic_sfm(24, n_landmarks=160, obs_per_camera=30, seed=7)          # the fixture's generator call
    Q, Abar = problems.q_from_observations(prob["N"], prob["M"], prob["cam"], prob["lm"], prob["w"], prob["pt"], return_abar=True)
    assert np.abs(Q - g["Q"]).max() <= 1e-12 * np.abs(g["Q"]).max()
    assert Abar.shape == g["Abar"].shape == (int(g["N"]) + int(g["M"]) - 1, 3 * int(g["N"]))
    assert np.abs(Abar - g["Abar"]).max() <= 1e-11 * np.abs(g["Abar"]).max()
    # Abar maps the noise-free optimum back to the ground-truth translations / landmarks (t_1 = 0 gauge)
    prob0 = problems.synthetic_sfm(24, n_landmarks=160, obs_per_camera=30, seed=7, noise=0.0)
    _, A0 = problems.q_from_observations(prob0["N"], prob0["M"], prob0["cam"], prob0["lm"], prob0["w"], prob0["pt"], return_abar=True)
    U = np.concatenate([prob0["s"][i] * prob0["R"][i].T for i in range(prob0["N"])], axis=0)     # 3N x 3 = (sR)^T stacked
    y = A0 @ U
    np.testing.assert_allclose(y[: prob0["N"] - 1], prob0["t"][1:], atol=1e-9)
    np.testing.assert_allclose(y[prob0["N"] - 1:], prob0["p"], atol=1e-9)
