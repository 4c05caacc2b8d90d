"""CPU, world_size 2 and 3 over gloo: the host logic of the row-sharded pass (partitioning, fixed-width
and CSR all-gathers, neighbour-list compaction) reassembles exactly what a single rank would hold."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, N, result_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import sys
        sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        from reid_gan_b200.sharded import RowComm, partition
        g = torch.Generator().manual_seed(123)
        # the "global truth" every rank can reconstruct
        rank_tbl = torch.randint(0, N, (N, 7), generator=g, dtype=torch.int32)
        cnt = torch.randint(0, 6, (N,), generator=g, dtype=torch.int32)
        cnt[::5] = 0
        ptr = torch.zeros(N + 1, dtype=torch.int64)
        ptr[1:] = torch.cumsum(cnt.to(torch.int64), 0)
        idx = torch.randint(0, N, (int(ptr[-1]),), generator=g, dtype=torch.int32)
        val = torch.rand(int(ptr[-1]), generator=g)
        comm = RowComm(N)
        r0, r1 = comm.r0, comm.r1
        assert (r0, r1) == partition(N, world, rank)
        # fixed-width rows
        got = comm.gather_rows(rank_tbl[r0:r1].clone())
        assert torch.equal(got, rank_tbl)
        # CSR pieces
        a, b = int(ptr[r0]), int(ptr[r1])
        g_ptr, g_idx, g_val, total, mx = comm.gather_csr(cnt[r0:r1].clone(), idx[a:b].clone(), val[a:b].clone())
        assert torch.equal(g_ptr, ptr) and total == int(ptr[-1]) and mx == int(cnt.max())
        assert torch.equal(g_idx[:total], idx) and torch.equal(g_val[:total], val)
        # neighbour lists stored in upper-bound slots (slot = 2 * count + 1 per row)
        n = r1 - r0
        slots = (2 * cnt[r0:r1].to(torch.int64) + 1)
        slot_ptr = torch.zeros(n + 1, dtype=torch.int64)
        slot_ptr[1:] = torch.cumsum(slots, 0)
        nbr = torch.full((int(slot_ptr[-1]),), -7, dtype=torch.int32)
        for i in range(n):
            c = int(cnt[r0 + i])
            nbr[int(slot_ptr[i]): int(slot_ptr[i]) + c] = idx[int(ptr[r0 + i]): int(ptr[r0 + i]) + c]
        n_ptr, n_idx, n_cnt = comm.gather_neighbors(slot_ptr, nbr, cnt[r0:r1].clone())
        assert torch.equal(n_ptr, ptr) and torch.equal(n_cnt, cnt) and torch.equal(n_idx[:total], idx)
        open(os.path.join(result_dir, "ok%d" % rank), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,N", [(2, 101), (3, 64), (2, 5)])
def test_row_comm_gloo(world, N, tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(world, port, N, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert (tmp_path / ("ok%d" % r)).exists()


def test_partition_covers_all_rows():
    from reid_gan_b200.sharded import partition
    for N in (1, 7, 32621, 250000):
        for W in (1, 2, 3, 4, 8):
            b = [partition(N, W, r) for r in range(W)]
            assert b[0][0] == 0 and b[-1][1] == N
            assert all(b[i][1] == b[i + 1][0] for i in range(W - 1))
            sizes = [y - x for x, y in b]
            assert max(sizes) - min(sizes) <= 1


def test_tile_dealing_is_balanced_and_complete():
    """Host logic of the tile-sharded search (sharded.knn_search_tiles): the upper-triangle tile list covers every
    (I <= J) exactly once and dealing tile (I, J) to rank (I + J) mod W spreads the tiles of every row block AND of
    every column block evenly, which is what keeps the per-rank partial candidate lists of a row equally long."""
    from reid_gan_b200 import knn_tc
    from reid_gan_b200.sharded import block_partition
    for n_t in (1, 7, 33, 128):
        t = knn_tc._tile_order(n_t, "cpu").numpy()
        assert t.shape == (n_t * (n_t + 1) // 2, 2)
        assert np.all(t[:, 0] <= t[:, 1])
        assert len({(int(a), int(b)) for a, b in t}) == t.shape[0]
        for W in (2, 3, 8):
            owner = (t[:, 0] + t[:, 1]) % W
            for blk in range(n_t):
                touching = owner[(t[:, 0] == blk) | (t[:, 1] == blk)]      # tiles that feed the rows of block `blk`
                counts = np.bincount(touching, minlength=W)
                assert counts.max() - counts.min() <= 1
    for N, W in ((32621, 8), (10, 4), (5, 8), (100000, 3)):
        blocks = [block_partition(N, W, r) for r in range(W)]
        B = blocks[0][2]
        assert all(b[2] == B for b in blocks) and blocks[0][0] == 0 and blocks[-1][1] == N
        assert all(blocks[r][1] == blocks[r + 1][0] for r in range(W - 1))
        assert all(0 <= b[1] - b[0] <= B for b in blocks)
        rows = np.arange(N)
        assert np.array_equal(np.minimum(rows // B, W - 1), np.concatenate([np.full(b[1] - b[0], r) for r, b in enumerate(blocks)]))


def test_shard_items_matches_the_row_partition():
    """f4: the share of the sorted training list a rank extracts features for is exactly the row block RowComm gives it
    (so `sharded.pseudo_labels(x_local, N=N)` receives the rows it expects), for every world size, with no row lost."""
    from reid_gan_b200.evaluators import shard_items
    from reid_gan_b200.sharded import partition
    items = [("img_%05d.jpg" % i, i % 7, i % 3) for i in range(1003)]
    for world in (1, 2, 3, 4, 8):
        seen = []
        for rank in range(world):
            share, r0, r1, n = shard_items(items, world=world, rank=rank)
            assert (r0, r1) == partition(len(items), world, rank) and n == len(items) and share == items[r0:r1]
            seen += share
        assert seen == items
