"""CPU: host-side bookkeeping of the tensor-core search (no kernels): tile orders, strip pairing, the sample-first
position tables, prepass splits."""
import numpy as np
import pytest
import torch


@pytest.mark.parametrize("n_t", [1, 2, 37, 128])
def test_tile_order_covers_the_upper_triangle_once(n_t):
    from reid_gan_b200 import knn_tc
    t = knn_tc._tile_order(n_t, "cpu").numpy()
    assert t.shape == (n_t * (n_t + 1) // 2, 2)
    assert (t[:, 0] <= t[:, 1]).all() and t.min() >= 0 and t.max() == n_t - 1
    assert len({(int(a), int(b)) for a, b in t}) == t.shape[0]


@pytest.mark.parametrize("n_t,first", [(33, 4), (128, 8), (9, 8)])
def test_tile_order_without_the_sample_blocks(n_t, first):
    from reid_gan_b200 import knn_tc
    t = knn_tc._tile_order_from(n_t, first, "cpu").numpy()
    full = knn_tc._tile_order(n_t, "cpu").numpy()
    want = {(int(a), int(b)) for a, b in full if a >= first}
    assert {(int(a), int(b)) for a, b in t} == want and t.shape[0] == len(want) == (n_t - first) * (n_t - first + 1) // 2


@pytest.mark.parametrize("sel", ["all", "chunk", "dealt"])
def test_strip_pairing_keeps_every_tile_once_and_in_order(sel):
    from reid_gan_b200 import knn_tc
    t = knn_tc._tile_order(64, "cpu")
    if sel == "chunk":                       # the tiles of one upload chunk (column blocks 20..24)
        t = t[(t[:, 1] >= 20) & (t[:, 1] < 25)]
    elif sel == "dealt":                     # the tiles dealt to rank 3 of 8
        t = t[((t[:, 0] + t[:, 1]) % 8) == 3]
    u = knn_tc.pair_units(t).numpy()
    flat = []
    for i, j0, j1 in u:
        flat.append((i, j0))
        if j1 >= 0:
            flat.append((i, j1))
    assert np.array_equal(np.array(flat, dtype=np.int32), t.numpy()), "order and multiplicity preserved"
    # a single tile only where the next tile belongs to another row block (or the list ends)
    k = 0
    for i, j0, j1 in u:
        k += 1 if j1 < 0 else 2
        if j1 < 0 and k < len(flat):
            assert flat[k][0] != i, "a tile is left single only at the end of its row block's run"
    assert knn_tc.pair_units(torch.empty((0, 2), dtype=torch.int32)).shape == (0, 3)


@pytest.mark.parametrize("N", [8269, 12936, 20480, 32621])
def test_sample_first_tables_are_inverse_permutations(N):
    from reid_gan_b200 import knn_tc
    m = knn_tc.sample_size(N)
    pos_of, orig_of = (t.numpy() for t in knn_tc._sample_first_maps(N, m, "cpu"))
    assert m % 256 == 0 and pos_of.shape == orig_of.shape == (N,)
    assert np.array_equal(np.sort(orig_of), np.arange(N)), "a permutation"
    assert np.array_equal(pos_of[orig_of], np.arange(N)) and np.array_equal(orig_of[pos_of], np.arange(N))
    stride = knn_tc._sample_stride(N, m)
    assert np.array_equal(orig_of[:m], (np.arange(m, dtype=np.int64) * stride) % N), "the low-discrepancy sample comes first"
    assert (np.diff(orig_of[m:]) > 0).all(), "the other rows keep their order"
    # the sample is spread over the whole row range (no block of N/16 rows without a sample row)
    gaps = np.diff(np.sort(orig_of[:m]))
    assert gaps.max() <= 2.7 * N / m, "three-gap bound of a golden-ratio walk"
    import math
    assert math.gcd(stride, N) == 1


def test_prepass_splits_fill_one_wave():
    from reid_gan_b200 import knn_tc
    assert knn_tc.prepass_splits(32621, 2048) == 1                       # 128 row units: no split
    assert knn_tc.prepass_splits(4096, 2048) == 4                        # 16 units x 4 splits = 64 <= 74 CTA-pair slots
    assert knn_tc.prepass_splits(8192, 2048) == 2                        # 32 units x 2
    assert knn_tc.prepass_splits(2048, 2048) == 4 and knn_tc.prepass_splits(2048, 2048, max_splits=8) == 8
    assert knn_tc.prepass_splits(2048, 1024, max_splits=8) == 4          # never more splits than sample tiles
