"""XM^2: the outer loop around the solver that the reference's pipeline scripts run (3_test_colmap_glomap.py:280-350) —
solve, evaluate every observation's residual, drop the worst 10 %, clean the graph, re-assemble Q, re-solve.

  observation_errors   w * || p_k - (s_i R_i p~_ik + t_i) ||^2 per observation, on the GPU (``xm_residuals``)
  outlier_cut          the reference's 90th-percentile rule (:318-327)
  check_landmarks      the graph clean-up of utils/checkconnection.py:15-100 (frames with <= 10 observations, landmarks seen
                       once, most-observed frame first, largest connected component) — restated with scipy.sparse.csgraph;
                       pinned by tests/golden/checklandmarks_ref.npz (outputs of the reference's own function)
  refine               the two-pass loop itself, on top of create_matrix / XM.solve / recover_XM

Host logic is NumPy like the reference's; everything numerical that touches the solver goes through the C-ABI (no CPU fallback).
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
from scipy.sparse.csgraph import connected_components


def _compact(counts_gt, size, idx0):
    """Index map old -> new (compact, order preserving) for entries observed more than `counts_gt` times; -1 otherwise."""
    cnt = np.bincount(idx0, minlength=size)
    keep = cnt > counts_gt
    new = np.full(size, -1, dtype=int)
    new[keep] = np.arange(int(keep.sum()))
    return int(np.argmax(cnt)), int(keep.sum()), new


def _compose(total_map, step_map):
    """total_map: original -> current (or -1); step_map: current -> next (or -1)."""
    out = total_map.copy()
    live = total_map > -1
    out[live] = step_map[total_map[live]]
    return out


def check_landmarks(edges, landmarks, weights, rgbs, N, M, min_frame_obs: int = 10, min_landmark_obs: int = 1):
    """Returns (edges, landmarks, weights, rgbs, frame_map) like utils/checkconnection.checklandmarks: edges stay 1-based
    (camera, landmark); frame_map[orig] = new 0-based camera index or -1."""
    edges = np.array(edges, dtype=int, copy=True)
    landmarks = np.asarray(landmarks); weights = np.asarray(weights); rgbs = np.asarray(rgbs)

    def drop(mask):
        nonlocal edges, landmarks, weights, rgbs
        edges = edges[~mask]; landmarks = landmarks[~mask]; weights = weights[~mask]; rgbs = rgbs[~mask]

    # 1. frames need more than `min_frame_obs` observations; the most observed one becomes camera 0 (the anchor)
    best, N, fmap = _compact(min_frame_obs, N, edges[:, 0] - 1)
    if fmap[best] != 0:
        fmap[fmap == 0] = fmap[best]
        fmap[best] = 0
    frame_map = fmap.copy()
    edges[:, 0] = fmap[edges[:, 0] - 1] + 1
    drop(np.any(edges == 0, axis=1))
    # 2. landmarks need more than `min_landmark_obs` observations
    _, M, lmap = _compact(min_landmark_obs, M, edges[:, 1] - 1)
    edges[:, 1] = lmap[edges[:, 1] - 1] + 1
    drop(np.any(edges == 0, axis=1))
    # 3. frames that lost all their landmarks
    _, N, fmap = _compact(0, N, edges[:, 0] - 1)
    edges[:, 0] = fmap[edges[:, 0] - 1] + 1
    frame_map = _compose(frame_map, fmap)
    drop(np.any(edges == 0, axis=1))
    # 4. largest connected component of the bipartite (camera, landmark) graph
    u = edges[:, 0] - 1; v = edges[:, 1] - 1 + N
    nn = N + M
    g = sp.coo_matrix((np.ones(u.size), (u, v)), shape=(nn, nn))
    ncomp, label = connected_components(g, directed=False)
    used = np.zeros(nn, dtype=bool); used[u] = True; used[v] = True
    sizes = np.bincount(label[used], minlength=ncomp)
    print("Number of connected components: ", int(np.count_nonzero(sizes)))
    # the reference takes max(components, key=len): the first of the largest in order of first appearance in the edge list
    cand = np.flatnonzero(sizes == sizes.max())
    if cand.size > 1:
        first = {c: i for i, c in reversed(list(enumerate(label[np.stack([u, v], axis=1).ravel()])))}
        big = min(cand, key=lambda c: first[c])
    else:
        big = cand[0]
    keep = label[u] == big
    if keep.sum() < edges.shape[0]:
        print("Not connected, Choose Largest Component")
        drop(~keep)
        _, N, fmap = _compact(0, N, edges[:, 0] - 1)
        edges[:, 0] = fmap[edges[:, 0] - 1] + 1
        frame_map = _compose(frame_map, fmap)
        _, M, lmap = _compact(0, M, edges[:, 1] - 1)
        edges[:, 1] = lmap[edges[:, 1] - 1] + 1
    return edges, landmarks, weights, rgbs, frame_map


def observation_errors(edges, landmarks, weights, R_real, s_real, t_est, p_est, handle=None):
    """Weighted squared residual of every observation (3_test_colmap_glomap.py:304-316), computed by ``xm_residuals``."""
    from . import capi
    h = handle or capi.Handle(device=0)
    edges = np.asarray(edges)
    return h.residuals(edges[:, 0] - 1, edges[:, 1] - 1, landmarks, weights, R_real, s_real, t_est, p_est)


def outlier_cut(errors, percentile: float = 90.0):
    """Indices of the observations to remove: error above the `percentile`-th percentile (:318-320)."""
    thr = np.percentile(errors, percentile)
    return np.where(errors > thr)[0]


def refine(edges, landmarks, weights, rgbs, N, M, output_path, solver=None, handle=None, max_rank: int = 5, tol: float = 1e-1,
           max_time: float = 1000.0):
    """The two-pass XM^2 loop of 3_test_colmap_glomap.py:280-350.  `solver` is the ``XM`` module (solve / solve_rank3);
    defaults to the compiled one.  Returns dict(R, s, p, t, edges, landmarks, weights, rgbs, frame_map, errors, lam)."""
    from . import binio
    from .creatematrix import create_matrix
    from .recover import recover_XM
    if solver is None:
        import XM as solver      # noqa: N811  (XM/build/ must be on sys.path, like in the reference's scripts)

    def load(name):
        return binio.load_matrix_from_bin(output_path + "/" + name)

    edges, landmarks, weights, rgbs, frame_map = check_landmarks(edges, landmarks, weights, rgbs, N, M)
    create_matrix(weights, edges, landmarks, output_path)
    lam = edges.shape[0] / N                 # the reference divides by the PRE-clean-up frame count (3_test_colmap_glomap.py:280-284)
    solver.solve(output_path, max_rank, tol, lam, max_time)
    R_real, s_real, p_est, t_est = recover_XM(load("Q.bin"), load("R.bin"), load("s.bin"), load("Abar.bin"), lam, handle=handle)
    errors = observation_errors(edges, landmarks, weights, R_real, s_real, t_est, p_est, handle=handle)
    print("sum of error: ", float(np.sum(errors)))
    rm = outlier_cut(errors)
    edges = np.delete(edges, rm, axis=0); weights = np.delete(weights, rm); rgbs = np.delete(rgbs, rm, axis=0)
    landmarks = np.delete(landmarks, rm, axis=0)
    # second run
    N = int(np.asarray(s_real).reshape(-1).shape[0])       # first-pass camera count (:296): also the divisor of the second lam (:341)
    M = int(p_est.shape[1])
    edges, landmarks, weights, rgbs, fmap2 = check_landmarks(edges, landmarks, weights, rgbs, N, M)
    frame_map = _compose(frame_map, fmap2)
    create_matrix(weights, edges, landmarks, output_path)
    lam = 0.0
    solver.solve_rank3(output_path, 3, tol, lam, max_time)
    s = load("s.bin")
    s_avg = np.mean(s[1:]); s_std = np.std(s[1:])
    if np.abs(s_avg - 1) > 2 * s_std or np.sum(s < 0.1) > 10:      # decide whether the scale regulariser is needed (:338-344)
        print("s is too small, run again")
        lam = edges.shape[0] / N
        solver.solve(output_path, max_rank, tol, lam, max_time)        # only in this branch (:339-342); otherwise the rank-3 result stands
    else:
        print("s is good")
    R_real, s_real, p_est, t_est = recover_XM(load("Q.bin"), load("R.bin"), load("s.bin"), load("Abar.bin"), lam, handle=handle)
    return dict(R=R_real, s=s_real, p=p_est, t=t_est, edges=edges, landmarks=landmarks, weights=weights, rgbs=rgbs,
                frame_map=frame_map, errors=errors, lam=lam)
