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
