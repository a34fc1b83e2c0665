"""Generates the committed fixtures under tests/golden/ from the REFERENCE (run in the build container only;
/root/reference does not exist on the GPU box).  Usage:  python tests/golden/make_goldens.py

  simple1_Q.bin            copy of the reference's shipped input assets/SIMPLE1/Q.bin (447 x 447, data fixture)
  simple2_obs.npz          SIMPLE2 observations after the preprocessing of 2_test_creatematrix.py:29-144
                           (edges 1-based like the reference, weights, camera-frame points) + ground-truth rotations
  simple2_frames.npz       original frame index of every re-indexed camera (to line gtR up with the solver's camera order)
  simple2_Q_ref.npz        Q (279 x 279) produced by the reference's own utils/creatematrix.create_matrix on them
  checklandmarks_ref.npz   inputs/outputs of the reference's own utils/checkconnection.checklandmarks on two random graphs
  recover_ref.npz          inputs/outputs of the reference's own utils/recoversolution.recover_XM (24 cameras, Q and Abar
                           from the reference's create_matrix): rank-3, rank-5 and mirrored cases
"""
import io
import os
import shutil
import sys
import tempfile
import contextlib

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REF)


def load_bin(fn):
    with open(fn, "rb") as f:
        r = int.from_bytes(f.read(4), "little"); c = int.from_bytes(f.read(4), "little")
        return np.fromfile(f, dtype=np.float64, count=r * c).reshape((r, c), order="F")


def preprocess(data):
    """The preprocessing steps of 2_test_creatematrix.py:31-144, restated (drop duplicate observations, re-index
    frames so the most-observed one is first, drop landmarks seen once, keep the largest connected component)."""
    import networkx as nx
    edges = data[:, :2].astype(int)
    _, uniq = np.unique(edges, axis=0, return_index=True)
    edges = edges[uniq]; data = data[uniq]
    weights = data[:, 5].copy(); pts = data[:, 2:5].copy()
    N = int(edges[:, 0].max()); M = int(edges[:, 1].max())

    def reindex(thr, size, idx0):
        cnt = np.bincount(idx0, minlength=size)
        valid = cnt > thr
        new = -np.ones(size, dtype=int); new[valid] = np.arange(valid.sum())
        return int(np.argmax(cnt)), int(valid.sum()), new

    maxf, N, fidx = reindex(0, N, edges[:, 0] - 1)
    if fidx[maxf] != 0:
        fidx[fidx == 0] = fidx[maxf]; fidx[maxf] = 0
    edges[:, 0] = fidx[edges[:, 0] - 1] + 1
    bad = np.any(edges == 0, axis=1)
    edges, weights, pts = edges[~bad], weights[~bad], pts[~bad]
    _, M, lidx = reindex(1, M, edges[:, 1] - 1)
    edges[:, 1] = lidx[edges[:, 1] - 1] + 1
    bad = np.any(edges == 0, axis=1)
    edges, weights, pts = edges[~bad], weights[~bad], pts[~bad]
    G = nx.Graph()
    G.add_edges_from((int(u), int(v) + N) for u, v in edges)
    comps = list(nx.connected_components(G))
    assert len(comps) == 1, "SIMPLE2 is connected; component filtering not needed"
    return edges, weights, pts, N, M


def frame_permutation(data):
    """orig_of_new[i] = original (0-based) frame index of re-indexed camera i: the preprocessing swaps the most-observed frame
    with frame 0 (2_test_creatematrix.py:74-110); ground truth files are in the original order."""
    edges = data[:, :2].astype(int)
    _, uniq = np.unique(edges, axis=0, return_index=True)
    edges = edges[uniq]
    N = int(edges[:, 0].max())
    cnt = np.bincount(edges[:, 0] - 1, minlength=N)
    assert np.all(cnt > 0)
    maxf = int(np.argmax(cnt))
    orig = np.arange(N)
    orig[0], orig[maxf] = maxf, 0
    return orig


def main():
    shutil.copyfile(f"{REF}/assets/SIMPLE1/Q.bin", f"{HERE}/simple1_Q.bin")
    data = load_bin(f"{REF}/assets/SIMPLE2/landmark.bin")
    edges, weights, pts, N, M = preprocess(data)
    gtR = load_bin(f"{REF}/assets/SIMPLE2/gtR.bin")
    np.savez(f"{HERE}/simple2_frames.npz", orig_of_new=frame_permutation(data))
    np.savez_compressed(f"{HERE}/simple2_obs.npz", edges=edges.astype(np.int32), weights=weights, pts=pts.astype(np.float64),
                        N=N, M=M, gtR=gtR)
    from utils.creatematrix import create_matrix
    tmp = tempfile.mkdtemp()
    with contextlib.redirect_stdout(io.StringIO()):
        create_matrix(weights, edges, pts, tmp)
    Q = load_bin(f"{tmp}/Q.bin")
    np.savez_compressed(f"{HERE}/simple2_Q_ref.npz", Q=Q)
    print("N", N, "M", M, "obs", edges.shape[0], "Q", Q.shape, "asym", np.abs(Q - Q.T).max())
    shutil.rmtree(tmp)


def make_recover_goldens():
    """recover_ref.npz: inputs and outputs of the reference's OWN utils/recoversolution.recover_XM on a small problem whose
    Q / Abar come from the reference's OWN create_matrix: (a) a rank-3 point, (b) a rank-5 point (top-3 eigen branch),
    (c) most cameras reflected (the global sign decision fires), (d) a few cameras reflected."""
    from utils.creatematrix import create_matrix
    from utils.recoversolution import recover_XM
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from xm_code_b200 import problems
    prob = problems.synthetic_sfm(24, n_landmarks=160, obs_per_camera=30, seed=7)
    N, M = prob["N"], prob["M"]
    edges = np.stack([prob["cam"] + 1, prob["lm"] + 1], axis=1).astype(int)
    tmp = tempfile.mkdtemp()
    with contextlib.redirect_stdout(io.StringIO()):
        create_matrix(prob["w"], edges, prob["pt"], tmp)
    Q = load_bin(f"{tmp}/Q.bin"); Abar = load_bin(f"{tmp}/Abar.bin")
    shutil.rmtree(tmp)
    rng = np.random.default_rng(5)
    out = dict(Q=Q, Abar=Abar, N=N, M=M)
    gt = np.concatenate([prob["R"][i] for i in range(N)], axis=0)      # 3N x 3, camera blocks stacked
    cases = {}
    # (a) near-ground-truth rank-3 point in an arbitrary gauge
    G, _ = np.linalg.qr(rng.standard_normal((3, 3)))
    cases["a"] = (gt @ G + 1e-3 * rng.standard_normal((3 * N, 3)), prob["s"].copy())
    # (b) rank 5: the same embedded in 5 columns by a random orthonormal frame + small off-subspace noise
    F, _ = np.linalg.qr(rng.standard_normal((5, 5)))
    R5 = np.concatenate([gt, np.zeros((3 * N, 2))], axis=1) @ F + 2e-3 * rng.standard_normal((3 * N, 5))
    cases["b"] = (R5, prob["s"] * rng.uniform(0.9, 1.1, N))
    # (c) most cameras (not the anchor) reflected -> "negative > N/2" global sign flip; (d) a minority reflected
    D = np.diag([1.0, 1.0, -1.0])
    Rc = gt.copy(); Rd = gt.copy()
    for i in range(1, N):
        if i % 3 != 0:
            Rc[3 * i:3 * i + 3] = Rc[3 * i:3 * i + 3] @ D
        if i % 5 == 1:
            Rd[3 * i:3 * i + 3] = Rd[3 * i:3 * i + 3] @ D
    cases["c"] = (Rc + 1e-3 * rng.standard_normal((3 * N, 3)), prob["s"].copy())
    cases["d"] = (Rd + 1e-3 * rng.standard_normal((3 * N, 3)), prob["s"].copy())
    for k, (R, s) in cases.items():
        with contextlib.redirect_stdout(io.StringIO()):
            R_real, s_real, p_est, t_est = recover_XM(Q, R, s[:, None], Abar, 0.0)
        out.update({f"{k}_R_in": R, f"{k}_s_in": s, f"{k}_R": R_real, f"{k}_s": s_real, f"{k}_p": p_est, f"{k}_t": t_est})
    np.savez_compressed(f"{HERE}/recover_ref.npz", **out)
    print("recover goldens: N", N, "M", M, "Abar", Abar.shape)




def make_checklandmarks_goldens():
    """checklandmarks_ref.npz: inputs and outputs of the reference's OWN utils/checkconnection.checklandmarks on two random
    bipartite graphs: (a) under-observed frames, landmarks seen once, the most-observed frame not first; (b) additionally two
    disconnected clusters (the largest-component branch)."""
    from utils.checkconnection import checklandmarks
    out = {}
    for tag, seed, split in (("a", 1, False), ("b", 2, True)):
        rng = np.random.default_rng(seed)
        N, M = 40, 300
        rows = []
        for i in range(N):
            k = int(rng.integers(3, 9)) if i % 7 == 3 else int(rng.integers(14, 40))     # some frames stay under the threshold of 10
            if i == 11:
                k = 80                                                                    # the most-observed frame is not frame 0
            if split:
                pool = np.arange(0, 200) if i < 28 else np.arange(200, 300)               # two clusters that share nothing
            else:
                pool = np.arange(M)
            for l in rng.choice(pool, size=min(k, pool.size), replace=False):
                rows.append((i + 1, int(l) + 1))
        edges = np.array(rows, dtype=int)
        n = edges.shape[0]
        landmarks = rng.standard_normal((n, 3)); weights = rng.uniform(0.5, 2.0, n); rgbs = rng.integers(0, 255, (n, 3))
        with contextlib.redirect_stdout(io.StringIO()):
            e2, l2, w2, c2, ind = checklandmarks(edges.copy(), landmarks.copy(), weights.copy(), rgbs.copy(), N, M)
        out.update({f"{tag}_edges_in": edges, f"{tag}_landmarks_in": landmarks, f"{tag}_weights_in": weights, f"{tag}_rgbs_in": rgbs,
                    f"{tag}_N": N, f"{tag}_M": M, f"{tag}_edges": e2, f"{tag}_landmarks": l2, f"{tag}_weights": w2, f"{tag}_rgbs": c2,
                    f"{tag}_indices": ind})
        print("checklandmarks golden", tag, edges.shape, "->", e2.shape, "frames", int(e2[:, 0].max()), "landmarks", int(e2[:, 1].max()))
    np.savez_compressed(f"{HERE}/checklandmarks_ref.npz", **out)


if __name__ == "__main__":
    if "--recover" in sys.argv:          # only the recover_XM fixture
        make_recover_goldens()
    elif "--checklandmarks" in sys.argv:
        make_checklandmarks_goldens()
    else:
        main()
        make_recover_goldens()
        make_checklandmarks_goldens()
