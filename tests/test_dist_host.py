"""CPU: host-side logic of the camera-partitioned (multi-GPU) solve — the partition arithmetic exported by the C-ABI,
slab extraction, and the handle exchange over a world_size-2 gloo process group.  No GPU compute here: NumPy stands in
for the per-rank product to check that the slabs tile the operator exactly."""
import os
import socket

import numpy as np
import pytest

from xm_code_b200 import capi, dist as xdist


@pytest.mark.parametrize("N,world,G", [(1723, 1, 148), (1723, 2, 148), (1778, 4, 148), (13682, 8, 148), (100000, 8, 148),
                                       (7, 2, 3), (300, 8, 37), (16, 8, 2)])
def test_partition_is_contiguous_balanced_and_cta_granular(N, world, G):
    tab = xdist.partition_table(N, world, G)
    assert tab[0][0] == 0 and tab[-1][1] == N
    for (a0, a1), (b0, b1) in zip(tab[:-1], tab[1:]):
        assert a1 == b0 and a0 <= a1
    sizes = [hi - lo for lo, hi in tab]
    assert max(sizes) - min(sizes) <= G            # each CTA owns floor or ceil of N / (world G) cameras
    GT = world * G
    for k, (lo, hi) in enumerate(tab):             # a rank's range is the union of its CTAs' ranges
        assert lo == (k * G * N) // GT and hi == ((k + 1) * G * N) // GT


def test_partition_rejects_bad_arguments():
    for args in [(0, 2, 148, 0), (10, 0, 148, 0), (10, 9, 148, 0), (10, 2, 0, 0), (10, 2, 148, 2), (10, 2, 148, -1)]:
        with pytest.raises(capi.XmError):
            capi.partition(*args)


def test_slabs_tile_dense_and_bsr_operators():
    rng = np.random.default_rng(0)
    N, world, G = 41, 4, 5
    Q = rng.standard_normal((3 * N, 3 * N)); X = rng.standard_normal((3 * N, 4))
    tab = xdist.partition_table(N, world, G)
    out = np.concatenate([xdist.row_slab(Q, lo, hi) @ X for lo, hi in tab])
    np.testing.assert_array_equal(out, np.concatenate([Q[3 * lo:3 * hi] @ X for lo, hi in tab]))
    assert out.shape == (3 * N, 4) and np.allclose(out, Q @ X)
    # block-CSR: random pattern, rows re-based per rank
    nnz_per_row = rng.integers(1, 6, N)
    rowptr = np.concatenate([[0], np.cumsum(nnz_per_row)]).astype(np.int32)
    col = np.concatenate([np.sort(rng.choice(N, k, replace=False)) for k in nnz_per_row]).astype(np.int32)
    vals = rng.standard_normal((rowptr[-1], 9))
    seen = 0
    for lo, hi in tab:
        rp, cc, vv = xdist.bsr_row_slab(rowptr, col, vals, lo, hi)
        assert rp[0] == 0 and rp[-1] == cc.shape[0] == vv.shape[0] and rp.size == hi - lo + 1
        np.testing.assert_array_equal(cc, col[rowptr[lo]:rowptr[hi]])
        seen += cc.shape[0]
    assert seen == rowptr[-1]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _gloo_worker(rank, world, port, N, G, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mine = bytes([rank + 1]) * capi.XM_IPC_HANDLE_BYTES
        allh = xdist.exchange_handles(mine)
        assert allh == [bytes([k + 1]) * capi.XM_IPC_HANDLE_BYTES for k in range(world)]
        # every rank computes the same table and its own slab product; the gathered result is the full product
        rng = np.random.default_rng(3)
        Q = rng.standard_normal((3 * N, 3 * N)); X = rng.standard_normal((3 * N, 3))
        lo, hi = xdist.partition_table(N, world, G)[rank]
        part = xdist.row_slab(Q, lo, hi) @ X
        parts = [None] * world
        dist.all_gather_object(parts, (lo, hi, part))
        full = np.concatenate([p[2] for p in sorted(parts, key=lambda t: t[0])])
        assert np.allclose(full, Q @ X)
        q.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_handle_exchange_and_sharding_over_gloo_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, 23, 4, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


def test_rcm_reorder_makes_the_contiguous_partition_a_graph_cut():
    """A banded view graph whose cameras arrive shuffled: after the RCM reorder the contiguous ranges need only a thin halo;
    an Erdos-Renyi graph needs (nearly) everything whatever the order.  The permuted operator is the same operator."""
    from xm_code_b200 import problems
    rng = np.random.default_rng(0)
    N = 600
    # band: camera i sees i-4..i+4, then a random relabelling hides the band
    rows = np.repeat(np.arange(N), 9); cols = (rows + np.tile(np.arange(-4, 5), N))
    ok = (cols >= 0) & (cols < N)
    rows, cols = rows[ok], cols[ok]
    relabel = rng.permutation(N)
    r2, c2 = relabel[rows], relabel[cols]
    order = np.lexsort((c2, r2)); r2, c2 = r2[order], c2[order]
    rowptr = np.concatenate([[0], np.cumsum(np.bincount(r2, minlength=N))]).astype(np.int32)
    col = c2.astype(np.int32)
    vals = rng.standard_normal((col.size, 3, 3))
    before = xdist.halo_statistics(rowptr, col, world=4, ctas_per_rank=10)
    perm = xdist.rcm_camera_order(rowptr, col)
    assert sorted(perm.tolist()) == list(range(N))
    rp2, col2, vals2 = xdist.permute_bsr(rowptr, col, vals, perm)
    after = xdist.halo_statistics(rp2, col2, world=4, ctas_per_rank=10)
    assert max(h["halo_fraction"] for h in before) > 0.5
    assert max(h["halo"] for h in after) <= 16 and max(h["halo_fraction"] for h in after) < 0.05
    # same operator: (P Q P^T)(P x) = P (Q x)
    Q = problems.bsr_to_dense(rowptr, col, np.ascontiguousarray(np.swapaxes(vals, 1, 2)))
    Q2 = problems.bsr_to_dense(rp2, col2, np.ascontiguousarray(np.swapaxes(vals2, 1, 2)))
    idx = (3 * perm[:, None] + np.arange(3)[None, :]).ravel()
    np.testing.assert_array_equal(Q2, Q[np.ix_(idx, idx)])
    # Erdos-Renyi: no cut to find
    rp, cc, vv = problems.erdos_renyi_bsr(2000, avg_degree=40, seed=1)
    er = xdist.halo_statistics(*xdist.permute_bsr(rp, cc, vv, xdist.rcm_camera_order(rp, cc))[:2], world=8, ctas_per_rank=10)
    assert min(h["halo_fraction"] for h in er) > 0.9
