"""GPU (-m gpu): the CUDA pass against the UNMODIFIED reference at the benchmark sizes.

tests/golden/full_{c1,c2,hard}.npz hold digests of the reference's own run (faiss_rerank.compute_jaccard_distance with
search_option=3 + sklearn DBSCAN on the dense matrix, generated in the build container by oracle/make_fullsize.py) on
  c1   N = 12,936  synth(n_ids=751)              BASELINE configs[0]
  c2   N = 32,621  synth(n_ids=1041)             BASELINE configs[1]  (the benchmark workload)
  hard N = 20,480  synth_hard: heavy-tailed identities, hub rows, duplicates, 2,380 DBSCAN noise points, 312 border points
Compared in full: every row of the neighbour lists (sha256 of initial_rank), every expansion set and every V_qe row
structure (counts + sha256 of the column ids), every eps-neighbourhood (counts + sha256), every label; a dozen complete
rows of J within 1e-5 with the identical ==1.0 pattern; centroid rows within 1e-4.  The same digests must come out of
row shards (virtual sharding, W = 2, 4, 8 on one GPU: SURVEY.md 8e "byte-identical for W=1,2,4,8")."""
import hashlib
import importlib
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = {
    "c1": ("synth", dict(N=12936, D=2048, n_ids=751, noise=0.8, seed=0)),
    "c2": ("synth", dict(N=32621, D=2048, n_ids=1041, noise=0.8, seed=0)),
    "hard": ("synth_hard", dict(N=20480, D=2048, seed=0)),
}


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def inputs(name):
    gen, kw = CASES[name]
    x, _ = getattr(importlib.import_module("reid_gan_b200.synth"), gen)(**kw)
    return x


def sorted_lists(ptr, idx, cnt):
    """Neighbour lists stored at ptr[i] .. ptr[i] + cnt[i] -> concatenation with every list sorted ascending."""
    ptr, idx, cnt = ptr.cpu().numpy(), idx.cpu().numpy(), cnt.cpu().numpy().astype(np.int64)
    rows = np.repeat(np.arange(cnt.size), cnt)
    first = np.cumsum(cnt) - cnt
    pos = ptr[:-1][rows] + (np.arange(rows.size) - first[rows]) if ptr.size == cnt.size + 1 else ptr[rows] + (np.arange(rows.size) - first[rows])
    flat = idx[pos]
    order = np.lexsort((flat, rows))
    return flat[order].astype(np.int32)


@pytest.mark.parametrize("name", ["c1", "hard", "c2"])
def test_full_size_against_reference_digest(name):
    import reid_gan_b200 as rg
    from reid_gan_b200.faiss_rerank import jaccard_neighbors
    g = np.load(os.path.join(GOLD, "full_%s.npz" % name))
    x = inputs(name)
    assert sha(x.numpy()) == str(g["x_sha256"]), "the seeded generator produced different bytes on this box"
    k1, k2, eps, ms = int(g["k1"]), int(g["k2"]), float(g["eps"]), int(g["min_samples"])
    N = x.shape[0]
    dist = rg.compute_jaccard_distance(x, k1=k1, k2=k2, print_flag=False, search_option=3)     # host features: streamed search
    st = dist.state
    # a1: every neighbour list
    assert sha(st.rank.cpu().numpy().astype(np.int32)) == str(g["rank_sha256"]), "initial_rank differs from the reference"
    # a3 / a5: every expansion set and every V_qe row (structure)
    e_cnt = (st.E_ptr[1:] - st.E_ptr[:-1]).cpu().numpy()
    assert np.array_equal(e_cnt, g["e_cnt"])
    assert sha(st.E_idx[: int(e_cnt.sum())].cpu().numpy().astype(np.int32)) == str(g["e_idx_sha256"])
    q_cnt = (st.Q_ptr[1:] - st.Q_ptr[:-1]).cpu().numpy()
    assert np.array_equal(q_cnt, g["q_cnt"])
    assert sha(st.Q_idx[: int(q_cnt.sum())].cpu().numpy().astype(np.int32)) == str(g["q_idx_sha256"])
    # a7: a dozen complete rows of the dense matrix (values within 1e-5, identical == 1.0 pattern)
    for r, ref in zip(g["j_rows"], g["j_vals"]):
        row = dist.dense_device(int(r), int(r) + 1).cpu().numpy()[0]
        assert np.array_equal(row == 1.0, ref == 1.0), "row %d: sparsity differs" % r
        assert np.abs(row - ref).max() <= 1e-5, "row %d: |dJ| = %g" % (r, np.abs(row - ref).max())
    # a7 + a8: every eps-neighbourhood and every label
    slot_ptr, nbr_idx, nbr_cnt, _ = jaccard_neighbors(st, eps)
    cnt = nbr_cnt.cpu().numpy()
    n_diff = int((cnt != g["nbr_cnt"]).sum())
    assert n_diff == 0, "%d rows have a different eps-neighbourhood size than the reference" % n_diff
    assert sha(sorted_lists(slot_ptr, nbr_idx, nbr_cnt)) == str(g["nbr_idx_sha256"]), "eps-neighbourhoods differ"
    assert bool(g["admissible"])
    labels = rg.DBSCAN(eps=eps, min_samples=ms, metric="precomputed", n_jobs=-1).fit_predict(dist)
    assert labels.dtype == np.intp and np.array_equal(labels, g["labels"]), "labels differ from sklearn on the reference's matrix"
    assert int((labels < 0).sum()) == int((g["labels"] < 0).sum())
    # a9
    cen = rg.generate_cluster_features(labels, x, normalize=True).cpu().numpy()
    assert cen.shape[0] == int(g["num_clusters"])
    np.testing.assert_allclose(cen[g["centroid_rows"]], g["centroid_vals"], rtol=1e-4, atol=1e-7)
    # the device-resident pass (what bench.py times) gives the same labels
    from reid_gan_b200 import pipeline
    out = pipeline.pseudo_labels(x.cuda(), k1, k2, eps, ms, centroids=True)
    assert np.array_equal(out["labels"].cpu().numpy(), g["labels"])
    np.testing.assert_allclose(out["centroids"].cpu().numpy()[g["centroid_rows"]], g["centroid_vals"], rtol=1e-4, atol=1e-7)


@pytest.mark.parametrize("name,W", [("hard", 2), ("hard", 4), ("hard", 8), ("c2", 8)])
def test_virtual_shards_are_byte_identical(name, W):
    """rerank_state(rows=(r0, r1)) for the W row blocks of a W-GPU run, on ONE GPU, must reproduce the
    unsharded state byte for byte: neighbour lists, keys, R masks, E / V rows, V_qe rows, eps-neighbourhoods."""
    import reid_gan_b200 as rg
    from reid_gan_b200.faiss_rerank import jaccard_neighbors, rerank_state
    from reid_gan_b200.sharded import partition
    g = np.load(os.path.join(GOLD, "full_%s.npz" % name))
    x = inputs(name).cuda()
    k1, k2, eps = int(g["k1"]), int(g["k2"]), float(g["eps"])
    N = x.shape[0]
    full = rerank_state(x, k1, k2)
    f_ptr, f_idx, f_cnt, f_val = jaccard_neighbors(full, eps, with_values=True)
    f_e = full.E_ptr.cpu().numpy()
    f_q = full.Q_ptr.cpu().numpy()
    f_rank, f_key = full.rank.cpu().numpy(), full.rank_key.cpu().numpy()
    f_lists = sorted_lists(f_ptr, f_idx, f_cnt)
    f_lptr = np.concatenate(([0], np.cumsum(f_cnt.cpu().numpy().astype(np.int64))))
    for r in range(W):
        r0, r1 = partition(N, W, r)
        st = rerank_state(x, k1, k2, rows=(r0, r1), knn_result=(full.rank, full.rank_key, dict(full.knn_info)))
        # own rows of every per-row stage
        assert np.array_equal(st.R_mask[: r1 - r0].cpu().numpy(), full.R_mask[r0:r1].cpu().numpy())
        e_ptr = st.E_ptr.cpu().numpy()
        assert np.array_equal(np.diff(e_ptr), np.diff(f_e[r0:r1 + 1]))
        ne = int(e_ptr[-1])
        assert np.array_equal(st.E_idx[:ne].cpu().numpy(), full.E_idx[int(f_e[r0]):int(f_e[r1])].cpu().numpy())
        assert np.array_equal(st.V_val[:ne].cpu().numpy().view(np.int32), full.V_val[int(f_e[r0]):int(f_e[r1])].cpu().numpy().view(np.int32))
        # the sharded search itself: own rows against all columns (one-sided tensor-core path + exact re-score)
        idx, key, _ = rg.knn_search(x, k1, "auto", rows=(r0, r1))
        assert np.array_equal(idx.cpu().numpy(), f_rank[r0:r1]) and np.array_equal(key.cpu().numpy().view(np.int32), f_key[r0:r1].view(np.int32))
    # V_qe rows and eps-neighbourhoods of a shard need the GLOBAL V / V_qe: take them from the full state, as the
    # all-gathers of sharded.py do, and recompute the shard's rows of a5 / a7
    from reid_gan_b200.faiss_rerank import query_expand_rows
    for r in range(W):
        r0, r1 = partition(N, W, r)
        q_cnt, q_idx, q_val = query_expand_rows(full, r0, r1)
        assert np.array_equal(q_cnt.cpu().numpy(), np.diff(f_q[r0:r1 + 1]))
        nq = int(q_cnt.sum())
        assert np.array_equal(q_idx[:nq].cpu().numpy(), full.Q_idx[int(f_q[r0]):int(f_q[r1])].cpu().numpy())
        assert np.array_equal(q_val[:nq].cpu().numpy().view(np.int32), full.Q_val[int(f_q[r0]):int(f_q[r1])].cpu().numpy().view(np.int32))
        full.row_begin, full.row_end = r0, r1
        try:
            s_ptr, s_idx, s_cnt, s_val = jaccard_neighbors(full, eps, with_values=True)
        finally:
            full.row_begin, full.row_end = 0, N
        assert np.array_equal(s_cnt.cpu().numpy(), f_cnt[r0:r1].cpu().numpy())
        assert np.array_equal(sorted_lists(s_ptr, s_idx, s_cnt), f_lists[f_lptr[r0]:f_lptr[r1]])


def test_speculative_sizes_that_do_not_hold_only_cost_time(monkeypatch):
    """The sync-free pass guesses the query-expansion table and the eps-neighbour storage; every kernel reports what
    did not fit and finish() redoes the pass with exact sizes.  Force both guesses to fail (and, separately, the kNN
    certificates): labels must still equal the reference's."""
    from reid_gan_b200 import faiss_rerank as fr, knn_tc, pipeline
    g = np.load(os.path.join(GOLD, "full_hard.npz"))
    x = inputs("hard").cuda()
    k1, k2, eps, ms = int(g["k1"]), int(g["k2"]), float(g["eps"]), int(g["min_samples"])
    out = pipeline.pseudo_labels(x, k1, k2, eps, ms)
    assert np.array_equal(out["labels"].cpu().numpy(), g["labels"])
    assert "speculation_failed" not in out["state"].knn_info, "the default guesses must hold on this set"
    # 1. eps-neighbour storage far too small
    monkeypatch.setattr(fr, "NBR_SPEC_PER_ROW", 1)
    fr._nbr_cap_hint.clear()
    out = pipeline.pseudo_labels(x, k1, k2, eps, ms)
    assert out["state"].report_vals[fr.R_NBR_OVF] > 0
    assert np.array_equal(out["labels"].cpu().numpy(), g["labels"])
    monkeypatch.undo()
    # 2. query-expansion table too small for most rows
    monkeypatch.setattr(fr, "QE_SPEC_SLOTS", 64)
    out = pipeline.pseudo_labels(x, k1, k2, eps, ms)
    assert out["state"].knn_info["speculation_failed"]["qe_overflow_rows"] > 0
    assert np.array_equal(out["labels"].cpu().numpy(), g["labels"])
    assert fr.rank_digest(out["state"]) == str(g["rank_sha256"])
    monkeypatch.undo()
    # 3. thresholds that leave fewer than k1 candidates: rows fail their certificate inside the deferred search
    monkeypatch.setattr(knn_tc, "SYM_TARGET", 16)
    out = pipeline.pseudo_labels(x, k1, k2, eps, ms)
    assert out["state"].knn_info["uncertified_rows"] > 0
    assert np.array_equal(out["labels"].cpu().numpy(), g["labels"])
    assert fr.rank_digest(out["state"]) == str(g["rank_sha256"])


@pytest.mark.parametrize("name", ["hard", "c1"])
def test_owned_pair_lists_are_the_full_lists_once(name):
    """The pass accumulates every unordered pair {i, j} in ONE of its two rows (J is bit-symmetric; csrc/jaccard.cu
    pair_owned).  Mirrored, the owned lists must be exactly the full eps-neighbourhoods (= the reference's, by the digest
    test above), no pair may be listed twice, and DBSCAN on them gives the same labels and core points."""
    from reid_gan_b200.dbscan import dbscan_from_neighbors
    from reid_gan_b200.faiss_rerank import jaccard_neighbors, rerank_state
    g = np.load(os.path.join(GOLD, "full_%s.npz" % name))
    x = inputs(name).cuda()
    k1, k2, eps, ms = int(g["k1"]), int(g["k2"]), float(g["eps"]), int(g["min_samples"])
    N = x.shape[0]
    st = rerank_state(x, k1, k2)
    f_ptr, f_idx, f_cnt, _ = jaccard_neighbors(st, eps)
    o_ptr, o_idx, o_cnt, _ = jaccard_neighbors(st, eps, owned=True)
    full = sorted_lists(f_ptr, f_idx, f_cnt)
    full_rows = np.repeat(np.arange(N), f_cnt.cpu().numpy().astype(np.int64))
    own = sorted_lists(o_ptr, o_idx, o_cnt).astype(np.int64)
    own_rows = np.repeat(np.arange(N), o_cnt.cpu().numpy().astype(np.int64))
    off = own != own_rows
    a = np.concatenate([own_rows, own[off]])
    b = np.concatenate([own, own_rows[off]])
    keys = a * N + b
    assert np.unique(keys).size == keys.size, "a pair is listed by both of its rows"
    assert np.array_equal(np.sort(keys), np.sort(full_rows * N + full.astype(np.int64)))
    assert abs(int(off.sum()) * 2 - int((full != full_rows).sum())) == 0
    lf, cf, nf = dbscan_from_neighbors(N, f_ptr, f_idx, f_cnt, ms)
    lo, co, no = dbscan_from_neighbors(N, o_ptr, o_idx, o_cnt, ms, owned=True)
    assert torch.equal(lf, lo) and torch.equal(cf, co) and int(nf) == int(no)
    assert np.array_equal(lo.cpu().numpy(), g["labels"])
