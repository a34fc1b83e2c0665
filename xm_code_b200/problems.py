"""Synthetic problem generators for the XM Burer-Monteiro path (no reference code involved).

The SDP data matrix follows the algebra of the reference's ``utils/creatematrix.py:52-341`` (verified against it on
SIMPLE2 in tests/test_problems.py via a committed golden):  for observations (camera i, landmark k, weight w,
camera-frame point p~) with cost  sum w || s_i R_i p~_ik + t_i - p_k ||^2 ,  eliminating translations and landmarks
with t_1 = 0 gives

    Q = Q1 - Vbar Lbar^{-1} Vbar^T ,   Q1 = blkdiag_i( sum_k w p~ p~^T ),   Vbar = [Vc  Vl](columns of t_1 removed),
    Lbar = weighted bipartite Laplacian of the (camera, landmark) graph without camera 1's row/column.

Landmarks are eliminated first (their block of Lbar is diagonal), leaving one dense (N-1) x (N-1) solve — this is
the sparse Schur assembly of SURVEY.md §8 row f2, and is what lets BAL-sized Q be built in seconds.
"""
from __future__ import annotations

import numpy as np
import scipy.linalg as sla
import scipy.sparse as sp


def q_from_observations(n_cameras: int, n_landmarks: int, cam, lm, w, pt, return_abar: bool = False):
    """Dense 3N x 3N float64 Q from observations.  cam, lm: int arrays (nobs); w: (nobs,); pt: (nobs, 3).
    return_abar: also return Abar ((N + M - 1) x 3N), the map (sR)^T -> [t_2..t_N; p_1..p_M] of the eliminated
    translations / landmarks (the reference's Abar.bin, utils/creatematrix.py:308-311): Abar = -Lbar^{-1} Vbar^T."""
    N, M = int(n_cameras), int(n_landmarks)
    cam = np.asarray(cam, dtype=np.int64); lm = np.asarray(lm, dtype=np.int64)
    w = np.asarray(w, dtype=np.float64); pt = np.asarray(pt, dtype=np.float64)
    nobs = cam.size
    # Q1: block diagonal sum w p p^T
    Q1 = np.zeros((N, 3, 3))
    np.add.at(Q1, cam, w[:, None, None] * pt[:, :, None] * pt[:, None, :])
    # Vc (3N x N): column i holds sum_k w p~ in camera i's rows; Vl (3N x M): -w p~ in camera rows, landmark column
    wp = w[:, None] * pt                                                  # (nobs, 3)
    Vc_blocks = np.zeros((N, 3)); np.add.at(Vc_blocks, cam, wp)
    rows = (3 * cam[:, None] + np.arange(3)[None, :]).ravel()
    cols = np.repeat(lm, 3)
    Vl = sp.csr_matrix((-wp.ravel(), (rows, cols)), shape=(3 * N, M))
    W = sp.csr_matrix((w, (cam, lm)), shape=(N, M))                      # camera-landmark weights
    dc = np.asarray(W.sum(axis=1)).ravel()                               # camera degrees
    dl = np.asarray(W.sum(axis=0)).ravel()                               # landmark degrees
    if np.any(dl <= 0):
        raise ValueError("every landmark needs at least one observation")
    Dl_inv = sp.diags(1.0 / dl)
    # eliminate landmarks: Vl Dl^-1 Vl^T (3N x 3N) and the reduced camera Laplacian
    VlD = Vl @ Dl_inv
    T1 = (VlD @ Vl.T).toarray()
    Sc = np.diag(dc) - (W @ Dl_inv @ W.T).toarray()                      # N x N reduced Laplacian
    # B = Vc + Vl Dl^-1 Wcl^T   (3N x N), then drop camera 0 (t_1 = 0)
    B = (VlD @ W.T).toarray()
    B[np.arange(3 * N), np.repeat(np.arange(N), 3)] += Vc_blocks.ravel()
    Bb = B[:, 1:]
    Scb = Sc[1:, 1:]
    cho = sla.cho_factor(Scb, lower=True, check_finite=False)
    T2 = Bb @ sla.cho_solve(cho, Bb.T, check_finite=False)
    Q = -(T1 + T2)
    idx = np.arange(N)
    for a in range(3):
        for b in range(3):
            Q[3 * idx + a, 3 * idx + b] += Q1[:, a, b]
    Q = 0.5 * (Q + Q.T)
    if not return_abar:
        return Q
    # Lbar [a; b] = -Vbar^T with the landmark block eliminated:  Sc a = -Bb^T ,  b = Dl^{-1} (-Vl^T + Wbar^T a)
    a_t = -sla.cho_solve(cho, Bb.T, check_finite=False)                  # (N-1) x 3N : translations t_2 .. t_N
    Wb = W[1:, :]
    b_p = Dl_inv @ (-Vl.T.toarray() + Wb.T @ a_t)                        # M x 3N     : landmarks
    return Q, np.vstack([a_t, b_p])


def random_rotations(n: int, rng) -> np.ndarray:
    A = rng.standard_normal((n, 3, 3))
    Qm, Rm = np.linalg.qr(A)
    Qm = Qm * np.sign(np.einsum("nii->ni", Rm))[:, None, :]
    det = np.linalg.det(Qm)
    Qm[:, :, 2] *= det[:, None]
    return Qm


def synthetic_sfm(n_cameras: int, n_landmarks: int | None = None, obs_per_camera: int = 40, turns: float = 2.0,
                  noise: float = 2e-3, seed: int = 0, **_ignored):
    """An object-centric synthetic capture shaped like the reference's shipped data (assets/SIMPLE2: narrow field of
    view, camera-frame depth ~1.15 +- 0.4): landmarks on a unit sphere (+ radial relief), cameras on a spiral of radius
    ~2.2 around it looking inward with random roll, each observing `obs_per_camera` landmarks on its near side.
    Consecutive cameras share landmarks (banded co-visibility graph, closed by the spiral's turns).
    Observation model of the reference (utils/creatematrix.py:52-98): s_i R_i p~_ik + t_i = p_k."""
    rng = np.random.default_rng(seed)
    N = int(n_cameras)
    M = int(n_landmarks) if n_landmarks else max(8 * N, 64)
    u = rng.standard_normal((M, 3)); u /= np.linalg.norm(u, axis=1, keepdims=True)
    p = u * (1.0 + 0.1 * rng.standard_normal((M, 1)))
    # camera centres on a spiral; world-from-camera rotation R_i has its +z axis pointing at the origin
    phi = np.linspace(0.0, 2.0 * np.pi * turns, N, endpoint=False) + 0.02 * rng.standard_normal(N)
    elev = 0.6 * np.sin(np.linspace(0.0, 2.0 * np.pi, N)) + 0.05 * rng.standard_normal(N)
    rad = 2.2 + 0.1 * rng.standard_normal(N)
    cpos = np.stack([rad * np.cos(elev) * np.cos(phi), rad * np.cos(elev) * np.sin(phi), rad * np.sin(elev)], axis=1)
    z = -cpos / np.linalg.norm(cpos, axis=1, keepdims=True)
    up = np.array([0.0, 0.0, 1.0])[None, :] + 0.15 * rng.standard_normal((N, 3))     # roughly upright cameras
    x = np.cross(up, z); x /= np.linalg.norm(x, axis=1, keepdims=True)
    y = np.cross(z, x)
    R = np.stack([x, y, z], axis=2)                      # columns = camera axes in the world frame
    s = np.concatenate([[1.0], rng.uniform(0.8, 1.25, N - 1)])
    t = cpos.copy()
    cam_l, lm_l = [], []
    for i in range(N):
        vis = np.nonzero(u @ (cpos[i] / np.linalg.norm(cpos[i])) > 0.55)[0]      # near-side cap, ~57 degrees
        k = min(obs_per_camera, vis.size)
        cam_l.append(np.full(k, i)); lm_l.append(rng.choice(vis, size=k, replace=False))
    cam = np.concatenate(cam_l); lm = np.concatenate(lm_l)
    cnt = np.bincount(lm, minlength=M)
    keep = cnt[lm] >= 2
    cam, lm = cam[keep], lm[keep]
    used = np.unique(lm)
    remap = -np.ones(M, dtype=np.int64); remap[used] = np.arange(used.size)
    lm = remap[lm]; p = p[used]; M = used.size
    # gauge of the reference: t_1 = 0 — shift the world so that camera 0 sits at the origin
    p = p - t[0]; t = t - t[0]
    pt = np.einsum("nba,nb->na", R[cam], p[lm] - t[cam]) / s[cam, None] + noise * rng.standard_normal((cam.size, 3))
    w = np.ones(cam.size)
    return dict(N=N, M=M, cam=cam, lm=lm, w=w, pt=pt, R=R, s=s, t=t, p=p)


def synthetic_dense_q(n_cameras: int, seed: int = 0, obs_per_camera: int = 40, noise: float = 2e-3, n_landmarks=None):
    """Dense Q of a synthetic problem with n_cameras cameras.  Returns (Q, problem dict with ground truth)."""
    prob = synthetic_sfm(n_cameras, n_landmarks=n_landmarks, obs_per_camera=obs_per_camera, noise=noise, seed=seed)
    Q = q_from_observations(prob["N"], prob["M"], prob["cam"], prob["lm"], prob["w"], prob["pt"])
    return Q, prob


def bal_shaped_q(name: str = "ladybug-1723", seed: int = 0):
    """Dense Q with the camera count of a BAL dataset (the BAL files themselves are not available offline).
    ladybug-1723: 3N = 5169 (213.7 MB); venice-1778: 3N = 5334."""
    sizes = {"ladybug-1723": 1723, "venice-1778": 1778, "final-13682": 13682}
    n = sizes[name]
    return synthetic_dense_q(n, seed=seed, obs_per_camera=60, n_landmarks=12 * n)


def erdos_renyi_bsr(n_cameras: int, avg_degree: float = 100.0, seed: int = 0, shared: int = 4):
    """Block-sparse PSD operator on an Erdos-Renyi view graph + Hamiltonian path (BASELINE config 5), 3x3 blocks.

    Each edge (i,j) carries `shared` synthetic point pairs; its cost sum_k || U_i a_k - U_j b_k ||^2 contributes
    A_i = sum a a^T to block (i,i), B_j = sum b b^T to (j,j) and -sum a b^T to (i,j) (transpose to (j,i)); a small
    ridge keeps every diagonal block well conditioned.  Returns (rowptr, colidx, vals[nnzb,3,3] column-major blocks)."""
    rng = np.random.default_rng(seed)
    N = int(n_cameras)
    m_target = int(avg_degree * N / 2)
    ii = rng.integers(0, N, size=m_target); jj = rng.integers(0, N, size=m_target)
    ok = ii != jj
    ii, jj = ii[ok], jj[ok]
    path_i = np.arange(N - 1); path_j = np.arange(1, N)
    ei = np.concatenate([path_i, np.minimum(ii, jj)]); ej = np.concatenate([path_j, np.maximum(ii, jj)])
    return _bsr_from_edges(N, ei, ej, rng, shared)


def banded_bsr(n_cameras: int, half_bandwidth: int = 6, seed: int = 0, shared: int = 4):
    """Block-sparse PSD operator on a BANDED view graph (camera i sees cameras i+1 .. i+half_bandwidth: a video-like capture) —
    the case where a graph-cut camera partition pays: contiguous camera ranges only share a thin boundary.  Same edge model and
    return convention as ``erdos_renyi_bsr``."""
    rng = np.random.default_rng(seed)
    N = int(n_cameras)
    ei = np.concatenate([np.arange(N - k) for k in range(1, half_bandwidth + 1)])
    ej = np.concatenate([np.arange(k, N) for k in range(1, half_bandwidth + 1)])
    return _bsr_from_edges(N, ei, ej, rng, shared)


def _bsr_from_edges(N, ei, ej, rng, shared):
    key = np.unique(np.minimum(ei, ej).astype(np.int64) * N + np.maximum(ei, ej))
    ei = (key // N).astype(np.int64); ej = (key % N).astype(np.int64)
    E = ei.size
    Rg = random_rotations(N, rng)
    pts = rng.standard_normal((E, shared, 3))
    a = np.einsum("eba,ekb->eka", Rg[ei], pts) + 0.01 * rng.standard_normal((E, shared, 3))
    b = np.einsum("eba,ekb->eka", Rg[ej], pts) + 0.01 * rng.standard_normal((E, shared, 3))
    Aii = np.einsum("eka,ekb->eab", a, a); Bjj = np.einsum("eka,ekb->eab", b, b); Cij = -np.einsum("eka,ekb->eab", a, b)
    diag = np.zeros((N, 3, 3))
    np.add.at(diag, ei, Aii); np.add.at(diag, ej, Bjj)
    diag += 1e-3 * np.eye(3)[None]
    rows = np.concatenate([np.arange(N), ei, ej]); cols = np.concatenate([np.arange(N), ej, ei])
    blocks = np.concatenate([diag, Cij, np.swapaxes(Cij, 1, 2)], axis=0)
    order = np.lexsort((cols, rows))
    rows, cols, blocks = rows[order], cols[order], blocks[order]
    rowptr = np.zeros(N + 1, dtype=np.int32); np.add.at(rowptr, rows + 1, 1); rowptr = np.cumsum(rowptr).astype(np.int32)
    vals = np.ascontiguousarray(np.swapaxes(blocks, 1, 2))     # per block column-major: vals[b, c, r] = block[r, c]
    return rowptr, cols.astype(np.int32), vals


def bsr_to_dense(rowptr, colidx, vals) -> np.ndarray:
    """Dense matrix of a 3x3 BSR operator (tests only; small N)."""
    N = rowptr.size - 1
    Q = np.zeros((3 * N, 3 * N))
    for i in range(N):
        for b in range(rowptr[i], rowptr[i + 1]):
            j = colidx[b]
            Q[3 * i:3 * i + 3, 3 * j:3 * j + 3] = vals[b].T
    return Q


# ------------------------------------------------------------------------------------------------ large problems (torch, on the GPU)
# BAL-Final-sized inputs (13 682 cameras, dense Q = 13.5 GB) take minutes with the NumPy/SciPy route above.  The two
# functions below restate the SAME generator and the SAME Schur assembly with torch tensor ops so they run on the
# device in seconds.  They are plain torch (index_put_, cholesky, matmul): input preparation shared by both arms of
# bench.py, not the product path (the product's own assembly is xm_create_matrix, xm_code_b200/csrc/xm_assemble.cu).
# Everything is deterministic for a given (seed, device type): accumulations go through index_put_(accumulate=True),
# which sorts on CUDA, and the selection of the observed landmarks uses an integer hash instead of a device RNG.

def _hash_u32(torch, a, b, seed):
    """Integer mixing of (a, b, seed) -> 32-bit keys (int64 tensors, broadcasting)."""
    m = 0xFFFFFFFF
    h = (a * 0x9E3779B1 + b * 0x85EBCA77 + (seed + 1) * 0xC2B2AE3D) & m
    h = h ^ (h >> 15)
    h = (h * 0x2C1B3C6D) & m
    h = h ^ (h >> 12)
    h = (h * 0x297A2D39) & m
    h = h ^ (h >> 15)
    return h


def synthetic_sfm_torch(n_cameras: int, n_landmarks: int | None = None, obs_per_camera: int = 60, turns: float = 2.0,
                        noise: float = 2e-3, seed: int = 0, device=None, chunk: int = 256):
    """The object-centric capture of ``synthetic_sfm`` (same geometry, same observation model) with the visibility test
    and the choice of the observed landmarks done by batched tensor ops on `device`.  Returns numpy arrays like
    ``synthetic_sfm`` (the observation lists are O(60 N): small)."""
    import torch
    dev = torch.device(device) if device is not None else torch.device("cuda" if torch.cuda.is_available() else "cpu")
    rng = np.random.default_rng(seed)
    N = int(n_cameras)
    M = int(n_landmarks) if n_landmarks else max(8 * N, 64)
    u = rng.standard_normal((M, 3)); u /= np.linalg.norm(u, axis=1, keepdims=True)
    p = u * (1.0 + 0.1 * rng.standard_normal((M, 1)))
    phi = np.linspace(0.0, 2.0 * np.pi * turns, N, endpoint=False) + 0.02 * rng.standard_normal(N)
    elev = 0.6 * np.sin(np.linspace(0.0, 2.0 * np.pi, N)) + 0.05 * rng.standard_normal(N)
    rad = 2.2 + 0.1 * rng.standard_normal(N)
    cpos = np.stack([rad * np.cos(elev) * np.cos(phi), rad * np.cos(elev) * np.sin(phi), rad * np.sin(elev)], axis=1)
    z = -cpos / np.linalg.norm(cpos, axis=1, keepdims=True)
    up = np.array([0.0, 0.0, 1.0])[None, :] + 0.15 * rng.standard_normal((N, 3))
    x = np.cross(up, z); x /= np.linalg.norm(x, axis=1, keepdims=True)
    y = np.cross(z, x)
    R = np.stack([x, y, z], axis=2)
    s = np.concatenate([[1.0], rng.uniform(0.8, 1.25, N - 1)])
    t = cpos.copy()
    # visibility (near-side cap) + hashed choice of obs_per_camera landmarks per camera, `chunk` cameras at a time
    ut = torch.from_numpy(u).to(dev)
    cn = torch.from_numpy(-z).to(dev)                    # unit vector from the origin to the camera
    lm_ids = torch.arange(M, device=dev, dtype=torch.int64)[None, :]
    k = min(obs_per_camera, M)
    BIG = float(1 << 40)
    cam_l, lm_l = [], []
    for c0 in range(0, N, chunk):
        c1 = min(N, c0 + chunk)
        vis = (cn[c0:c1] @ ut.T) > 0.55
        ci = torch.arange(c0, c1, device=dev, dtype=torch.int64)[:, None]
        key = _hash_u32(torch, ci, lm_ids, seed).to(torch.float64)
        key = torch.where(vis, key, torch.full_like(key, BIG))
        val, idx = torch.topk(key, k, dim=1, largest=False)
        ok = val < BIG
        cam_l.append(ci.expand(-1, k)[ok]); lm_l.append(idx[ok])
    cam = torch.cat(cam_l); lm = torch.cat(lm_l)
    cnt = torch.bincount(lm, minlength=M)
    keep = cnt[lm] >= 2
    cam, lm = cam[keep].cpu().numpy(), lm[keep].cpu().numpy()
    order = np.lexsort((lm, cam))
    cam, lm = cam[order], lm[order]
    used = np.unique(lm)
    remap = -np.ones(M, dtype=np.int64); remap[used] = np.arange(used.size)
    lm = remap[lm]; p = p[used]; M = used.size
    p = p - t[0]; t = t - t[0]
    pt = np.einsum("nba,nb->na", R[cam], p[lm] - t[cam]) / s[cam, None] + noise * rng.standard_normal((cam.size, 3))
    w = np.ones(cam.size)
    return dict(N=N, M=M, cam=cam, lm=lm, w=w, pt=pt, R=R, s=s, t=t, p=p)


def q_from_observations_torch(n_cameras: int, n_landmarks: int, cam, lm, w, pt, device=None):
    """``q_from_observations`` on `device` with torch: returns the dense 3N x 3N float64 Q as a torch tensor (symmetric, so
    its memory is both the row-major and the column-major matrix).  Peak memory ~ 2.7 x the matrix."""
    import torch
    dev = torch.device(device) if device is not None else torch.device("cuda" if torch.cuda.is_available() else "cpu")
    N, M = int(n_cameras), int(n_landmarks)
    f64 = torch.float64
    cam_t = torch.as_tensor(np.asarray(cam, dtype=np.int64), device=dev)
    lm_t = torch.as_tensor(np.asarray(lm, dtype=np.int64), device=dev)
    w_t = torch.as_tensor(np.asarray(w, dtype=np.float64), device=dev)
    pt_t = torch.as_tensor(np.asarray(pt, dtype=np.float64), device=dev)
    nobs = cam_t.numel()
    n3 = 3 * N
    wp = w_t[:, None] * pt_t                                              # (nobs, 3)
    dl = torch.zeros(M, dtype=f64, device=dev).index_put_((lm_t,), w_t, accumulate=True)
    dc = torch.zeros(N, dtype=f64, device=dev).index_put_((cam_t,), w_t, accumulate=True)
    if bool((dl <= 0).any()):
        raise ValueError("every landmark needs at least one observation")
    # ordered pairs (a, b) of observations of the same landmark
    order = torch.argsort(lm_t, stable=True)
    lm_s = lm_t[order]
    deg = torch.bincount(lm_s, minlength=M)
    start = torch.cumsum(deg, 0) - deg                                    # first sorted position of each landmark
    d_a = deg[lm_s]                                                       # group size of each sorted observation
    a_pos = torch.repeat_interleave(torch.arange(nobs, device=dev), d_a)
    off = torch.cumsum(d_a, 0) - d_a
    b_pos = start[lm_s][a_pos] + (torch.arange(a_pos.numel(), device=dev) - off[a_pos])
    a = order[a_pos]; b = order[b_pos]
    inv_dl = 1.0 / dl[lm_t[a]]
    ca, cb = cam_t[a], cam_t[b]
    three = torch.arange(3, device=dev)
    Q = torch.zeros((n3, n3), dtype=f64, device=dev)
    # T1 = Vl Dl^-1 Vl^T : block (ca, cb) += wp_a wp_b^T / dl ;  Q starts as -T1
    blk = -(wp[a][:, :, None] * wp[b][:, None, :]) * inv_dl[:, None, None]
    rows = (3 * ca)[:, None, None] + three[None, :, None]
    cols = (3 * cb)[:, None, None] + three[None, None, :]
    Q.view(-1).index_put_(((rows * n3 + cols).reshape(-1),), blk.reshape(-1), accumulate=True)
    del blk, rows, cols
    # Q1 on the diagonal blocks
    q1 = w_t[:, None, None] * pt_t[:, :, None] * pt_t[:, None, :]
    rows = (3 * cam_t)[:, None, None] + three[None, :, None]
    cols = (3 * cam_t)[:, None, None] + three[None, None, :]
    Q.view(-1).index_put_(((rows * n3 + cols).reshape(-1),), q1.reshape(-1), accumulate=True)
    del q1, rows, cols
    # reduced camera Laplacian Sc = diag(dc) - W Dl^-1 W^T and B = Vc + Vl Dl^-1 W^T
    Sc = torch.zeros((N, N), dtype=f64, device=dev)
    Sc.view(-1).index_put_((ca * N + cb,), -(w_t[a] * w_t[b] * inv_dl), accumulate=True)
    Sc.diagonal().add_(dc)
    B = torch.zeros((n3, N), dtype=f64, device=dev)
    vals = -(wp[a] * (w_t[b] * inv_dl)[:, None])                          # (pairs, 3)
    rows = (3 * ca)[:, None] + three[None, :]
    B.view(-1).index_put_(((rows * N + cb[:, None]).reshape(-1),), vals.reshape(-1), accumulate=True)
    rows = (3 * cam_t)[:, None] + three[None, :]
    B.view(-1).index_put_(((rows * N + cam_t[:, None]).reshape(-1),), wp.reshape(-1), accumulate=True)
    del vals, rows, a, b, ca, cb, a_pos, b_pos
    Bb = B[:, 1:].contiguous(); del B
    L = torch.linalg.cholesky(Sc[1:, 1:]); del Sc
    Z = torch.cholesky_solve(Bb.T.contiguous(), L); del L                  # (N-1) x 3N
    Q.addmm_(Bb, Z, alpha=-1.0); del Bb, Z
    Q = (Q + Q.T).mul_(0.5)
    return Q


def synthetic_dense_q_torch(n_cameras: int, seed: int = 0, obs_per_camera: int = 60, n_landmarks=None, noise: float = 2e-3, device=None):
    """Device-side twin of ``synthetic_dense_q`` (hashed landmark choice: a different, equally shaped problem)."""
    prob = synthetic_sfm_torch(n_cameras, n_landmarks=n_landmarks, obs_per_camera=obs_per_camera, noise=noise, seed=seed, device=device)
    Q = q_from_observations_torch(prob["N"], prob["M"], prob["cam"], prob["lm"], prob["w"], prob["pt"], device=device)
    return Q, prob
