"""Host-side plumbing for ONE solve partitioned by camera over the GPUs of a node.

One process per GPU; ``torch.distributed`` (nccl on the GPU box, gloo in the CPU tests) is used ONLY to hand the 64-byte
CUDA-IPC handles of the ranks' exchange arenas around and to line the ranks up.  The data path — the per-iteration
all-gather of the Q.Y operand rows, the scalar all-reduces and the barriers — is peer-mapped stores issued from inside
the persistent kernels (include/xm_b200.h, "multi-GPU"); no collective library call sits on it.

The reference is single-GPU (XM/include/Utils/memory.h:284,366: ``gpu_id = 0`` everywhere), so there is no reference
interface to mirror here; the calls below keep the C-ABI's names.
"""
from __future__ import annotations

import numpy as np

from . import capi


def partition_table(n_cameras: int, world: int, ctas_per_rank: int):
    """[(cam_lo, cam_hi)] per rank — contiguous, disjoint, covering range(n_cameras)."""
    return [capi.partition(n_cameras, world, ctas_per_rank, k) for k in range(world)]


def row_slab(Q: np.ndarray, cam_lo: int, cam_hi: int) -> np.ndarray:
    """Rows of the dense 3N x 3N matrix that a rank owning cameras [cam_lo, cam_hi) uploads (a view, no copy)."""
    return Q[3 * cam_lo:3 * cam_hi, :]


def bsr_row_slab(rowptr, colidx, vals, cam_lo: int, cam_hi: int):
    """Block rows [cam_lo, cam_hi) of a block-CSR matrix, re-based (what xm_set_q_bsr keeps on a rank)."""
    rowptr = np.asarray(rowptr)
    b0, b1 = int(rowptr[cam_lo]), int(rowptr[cam_hi])
    return (rowptr[cam_lo:cam_hi + 1] - b0).astype(np.int32), np.asarray(colidx)[b0:b1], np.asarray(vals)[b0:b1]


def exchange_handles(local: bytes, group=None):
    """all_gather of every rank's IPC handle, in rank order (works on gloo and nccl process groups)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    mine = torch.frombuffer(bytearray(local), dtype=torch.uint8).to(dev)
    out = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(out, mine, group=group)
    return [bytes(t.cpu().numpy().tobytes()) for t in out]


def attach(handle: "capi.Handle", n_cameras: int, max_r: int, group=None) -> dict:
    """Turn `handle` into this rank's member of a communicator spanning the process group.  Collective."""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if world > capi.XM_MAX_WORLD:
        raise capi.XmError(f"at most {capi.XM_MAX_WORLD} GPUs of one node can share a solve")
    mine = handle.comm_init(rank, world, n_cameras, max_r)
    handles = exchange_handles(mine, group)
    info = handle.comm_info()
    # every rank must have planned the same number of CTAs (same GPU model): the partition depends on it
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    g = torch.tensor([info["ctas_per_rank"]], dtype=torch.int64, device=dev)
    gs = [torch.empty_like(g) for _ in range(world)]
    dist.all_gather(gs, g, group=group)
    if len({int(t.item()) for t in gs}) != 1:
        raise capi.XmError(f"ranks disagree on CTAs per rank: {[int(t.item()) for t in gs]}")
    handle.comm_connect(handles)
    dist.barrier(group)          # nobody launches before every rank has mapped every arena
    return info


def reset(handle: "capi.Handle", group=None):
    """Recover a communicator after XM_ESYNC (collective)."""
    import torch.distributed as dist
    dist.barrier(group)
    handle.comm_reset()
    dist.barrier(group)


def detach(handle: "capi.Handle", group=None):
    """Leave the communicator (collective): every rank unmaps its peers' arenas before anybody frees its own."""
    import torch.distributed as dist
    handle.comm_disconnect()
    dist.barrier(group)
    handle.close()


# ------------------------------------------------------------------------------------------------ graph-cut partition (block-CSR)
def rcm_camera_order(rowptr, colidx):
    """Reverse Cuthill-McKee order of the cameras of a block-CSR view graph ("simple BFS/RCM bands"): cameras
    that see each other end up close together, so the library's contiguous camera ranges become a band partition with few
    boundary cameras.  Returns perm with perm[new] = old."""
    return capi.rcm_order(rowptr, colidx)          # xm_rcm_order (host C++ in the library)


def permute_bsr(rowptr, colidx, vals, perm):
    """Symmetric permutation P Q P^T of a block-CSR matrix: new camera k is old camera perm[k].  Block values are carried
    along unchanged (rows and columns of a 3x3 block belong to one camera each)."""
    rowptr = np.asarray(rowptr); colidx = np.asarray(colidx); vals = np.asarray(vals)
    n = rowptr.size - 1
    inv = np.empty(n, dtype=np.int64); inv[perm] = np.arange(n)
    counts = np.diff(rowptr)[perm]
    new_rowptr = np.concatenate([[0], np.cumsum(counts)]).astype(rowptr.dtype)
    src = np.concatenate([np.arange(rowptr[o], rowptr[o + 1]) for o in perm]) if n else np.zeros(0, dtype=np.int64)
    new_col = inv[colidx[src]]
    # sort every row by its new column index (the library does not need it, CSR consumers usually expect it)
    row_of = np.repeat(np.arange(n), counts)
    order = np.lexsort((new_col, row_of))
    return new_rowptr, new_col[order].astype(colidx.dtype), vals[src][order]


def halo_statistics(rowptr, colidx, world: int, ctas_per_rank: int = 148):
    """For the library's contiguous camera partition: per rank, how many cameras it owns, how many REMOTE cameras its block
    rows reference (the halo a boundary-only exchange would have to fetch) and the fraction of all remote cameras that is —
    1.0 means the full all-gather the kernels do today is already minimal (Erdos-Renyi), << 1 means a boundary-only exchange
    would move that much less (banded view graphs after rcm_camera_order)."""
    rowptr = np.asarray(rowptr); colidx = np.asarray(colidx)
    n = rowptr.size - 1
    out = []
    for k, (lo, hi) in enumerate(partition_table(n, world, ctas_per_rank)):
        cols = np.unique(colidx[rowptr[lo]:rowptr[hi]])
        remote = cols[(cols < lo) | (cols >= hi)]
        out.append(dict(rank=k, cameras=hi - lo, halo=int(remote.size), remote_total=n - (hi - lo),
                        halo_fraction=float(remote.size) / max(1, n - (hi - lo))))
    return out
