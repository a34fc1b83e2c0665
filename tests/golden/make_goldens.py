"""Generates the committed fixtures under tests/golden/ from the REFERENCE (run in the build container only;
/root/reference does not exist on the GPU box).  Usage:  python tests/golden/make_goldens.py

  simple1_Q.bin            copy of the reference's shipped input assets/SIMPLE1/Q.bin (447 x 447, data fixture)
  simple2_obs.npz          SIMPLE2 observations after the preprocessing of 2_test_creatematrix.py:29-144
                           (edges 1-based like the reference, weights, camera-frame points) + ground-truth rotations
  simple2_Q_ref.npz        Q (279 x 279) produced by the reference's own utils/creatematrix.create_matrix on them
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


def main():
    shutil.copyfile(f"{REF}/assets/SIMPLE1/Q.bin", f"{HERE}/simple1_Q.bin")
    data = load_bin(f"{REF}/assets/SIMPLE2/landmark.bin")
    edges, weights, pts, N, M = preprocess(data)
    gtR = load_bin(f"{REF}/assets/SIMPLE2/gtR.bin")
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


if __name__ == "__main__":
    main()
