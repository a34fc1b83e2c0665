"""Q assembly on the GPU (xm_create_matrix) against OUTPUT OF THE REFERENCE'S OWN create_matrix
(utils/creatematrix.py:52-341; goldens made by tests/golden/make_goldens.py): Q on SIMPLE2 (1e-10), Q and Abar on the 24-camera
fixture (1e-13 relative on Abar), plus the host restatement on a larger synthetic problem and the plain-torch generator."""
import os

import numpy as np
import pytest

from conftest import GOLD
from xm_code_b200 import problems

pytestmark = pytest.mark.gpu


def test_simple2_q_matches_the_reference(gpu_handle_factory, simple2_obs, simple2_q):
    o = simple2_obs
    N, M = int(o["N"]), int(o["M"])
    h = gpu_handle_factory()
    Q, Abar, ms = h.create_matrix(N, M, o["edges"][:, 0] - 1, o["edges"][:, 1] - 1, o["weights"], o["pts"], want_q=True, want_abar=True)
    scale = np.abs(simple2_q).max()
    assert np.abs(Q - simple2_q).max() <= 1e-10 * scale
    assert np.array_equal(Q, Q.T) and Abar.shape == (N + M - 1, 3 * N) and ms > 0
    # the assembled matrix IS the handle's operator now: a product needs no upload
    X = np.random.default_rng(0).standard_normal((3 * N, 3))
    assert np.abs(h.qy(X) - simple2_q @ X).max() <= 1e-9 * np.abs(simple2_q @ X).max()
    Qh, Abar_h = problems.q_from_observations(N, M, o["edges"][:, 0] - 1, o["edges"][:, 1] - 1, o["weights"], o["pts"], return_abar=True)
    assert np.abs(Abar - Abar_h).max() <= 1e-10 * np.abs(Abar_h).max()


def test_24_camera_fixture_q_and_abar_match_the_reference(gpu_handle_factory):
    g = np.load(os.path.join(GOLD, "recover_ref.npz"))
    prob = problems.synthetic_sfm(24, n_landmarks=160, obs_per_camera=30, seed=7)          # the fixture's generator call (make_goldens.py)
    N, M = prob["N"], prob["M"]
    assert N == int(g["N"]) and M == int(g["M"])
    h = gpu_handle_factory()
    Q, Abar, _ = h.create_matrix(N, M, prob["cam"], prob["lm"], prob["w"], prob["pt"], want_q=True, want_abar=True)
    assert np.abs(Q - g["Q"]).max() <= 1e-10 * np.abs(g["Q"]).max()
    assert np.abs(Abar - g["Abar"]).max() <= 1e-10 * np.abs(g["Abar"]).max()


def test_create_matrix_drop_in_writes_the_reference_files(tmp_path):
    """creatematrix.create_matrix(weight, edges, landmarks, output_path): the reference's signature, 1-based edges, Q.bin / Abar.bin."""
    from xm_code_b200 import binio, creatematrix
    g = np.load(os.path.join(GOLD, "recover_ref.npz"))
    prob = problems.synthetic_sfm(24, n_landmarks=160, obs_per_camera=30, seed=7)
    edges = np.stack([prob["cam"] + 1, prob["lm"] + 1], axis=1)
    Q, Abar = creatematrix.create_matrix(prob["w"], edges, prob["pt"], str(tmp_path))
    assert np.abs(Q - g["Q"]).max() <= 1e-10 * np.abs(g["Q"]).max()
    np.testing.assert_array_equal(binio.load_matrix_from_bin(str(tmp_path / "Q.bin")), Q)
    np.testing.assert_array_equal(binio.load_matrix_from_bin(str(tmp_path / "Abar.bin")), Abar)


def test_larger_problem_matches_host_restatement_and_torch_generator(gpu_handle_factory):
    prob = problems.synthetic_sfm(400, n_landmarks=4800, obs_per_camera=60, seed=2)
    h = gpu_handle_factory()
    Q, _, ms = h.create_matrix(prob["N"], prob["M"], prob["cam"], prob["lm"], prob["w"], prob["pt"], want_q=True, want_abar=False)
    Qh = problems.q_from_observations(prob["N"], prob["M"], prob["cam"], prob["lm"], prob["w"], prob["pt"])
    assert np.abs(Q - Qh).max() <= 1e-11 * np.abs(Qh).max()
    Qt = problems.q_from_observations_torch(prob["N"], prob["M"], prob["cam"], prob["lm"], prob["w"], prob["pt"], device="cuda").cpu().numpy()
    assert np.abs(Qt - Qh).max() <= 1e-11 * np.abs(Qh).max()
    # duplicate observations of one landmark by one camera and non-unit weights
    rng = np.random.default_rng(5)
    cam = np.concatenate([prob["cam"], prob["cam"][:50]]); lm = np.concatenate([prob["lm"], prob["lm"][:50]])
    w = rng.uniform(0.5, 2.0, cam.size); pt = np.concatenate([prob["pt"], prob["pt"][:50] + 0.01])
    Q2, _, _ = h.create_matrix(prob["N"], prob["M"], cam, lm, w, pt, want_q=True)
    Q2h = problems.q_from_observations(prob["N"], prob["M"], cam, lm, w, pt)
    assert np.abs(Q2 - Q2h).max() <= 1e-11 * np.abs(Q2h).max()


def test_bad_inputs_are_rejected(gpu_handle_factory):
    from xm_code_b200 import capi
    h = gpu_handle_factory()
    with pytest.raises(capi.XmError):       # landmark 3 never observed
        h.create_matrix(3, 4, [0, 1, 2, 0, 1, 2], [0, 0, 1, 1, 2, 2], np.ones(6), np.ones((6, 3)))
    with pytest.raises(capi.XmError):       # camera index out of range
        h.create_matrix(3, 2, [0, 1, 5], [0, 0, 1], np.ones(3), np.ones((3, 3)))
