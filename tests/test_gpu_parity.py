"""GPU (-m gpu): the CUDA path, called through the C ABI, against the golden vectors of the
reference and against the oracle on seeded inputs.  Bars (BASELINE.json north_star):
neighbour lists / reciprocal sets / expansion sets / DBSCAN labels bit-exact; Jaccard distance
within 1e-5 absolute; centroids within 1e-4 relative."""
import glob
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
RERANK = sorted(glob.glob(os.path.join(GOLD, "rerank_*.npz")))
KNN_MODES = ("exact", "auto")
J_ATOL = 1e-5


def golden_J(g):
    N = g["x"].shape[0]
    J = np.ones((N, N), dtype=np.float32)
    J[g["J_rows"], g["J_cols"]] = g["J_vals"]
    return J


def masks_to_lists(rank, mask):
    out = []
    for i in range(rank.shape[0]):
        m = int(mask[i]) & 0xFFFFFFFFFFFFFFFF
        out.append(np.array([rank[i, r] for r in range(rank.shape[1]) if (m >> r) & 1], dtype=np.int64))
    return out


@pytest.mark.parametrize("knn", KNN_MODES)
@pytest.mark.parametrize("path", RERANK, ids=[os.path.basename(p)[:-4] for p in RERANK])
def test_golden_jaccard_and_labels(path, knn):
    import reid_gan_b200 as rg
    g = np.load(path)
    x, k1, k2 = torch.from_numpy(g["x"]), int(g["k1"]), int(g["k2"])
    dist = rg.compute_jaccard_distance(x, k1=k1, k2=k2, print_flag=False, search_option=3, knn=knn)
    J = np.asarray(dist)
    J_ref = golden_J(g)
    assert J.dtype == np.float32 and J.shape == J_ref.shape
    assert np.array_equal(J == 1.0, J_ref == 1.0), "sparsity structure differs from the reference"
    assert np.abs(J - J_ref).max() <= J_ATOL
    assert np.array_equal(J, J.T), "J must be bit-symmetric like the reference's"
    for eps in g["eps_list"]:
        tag = "eps%02d" % round(float(eps) * 100)
        if not bool(g["admissible_" + tag]):
            continue
        c = rg.DBSCAN(eps=float(eps), min_samples=4, metric="precomputed", n_jobs=-1)
        assert np.array_equal(c.fit_predict(dist), g["labels_" + tag]), "sparse-path labels"
        assert c.labels_.dtype == np.intp
        assert np.array_equal(rg.DBSCAN(eps=float(eps), min_samples=4).fit_predict(J_ref), g["labels_" + tag]), \
            "dense-path labels on the reference's own matrix"
        if "centroids_" + tag in g.files:
            cen = rg.generate_cluster_features(g["labels_" + tag], x, normalize=True).cpu().numpy()
            np.testing.assert_allclose(cen, g["centroids_" + tag], rtol=1e-4, atol=1e-7)


@pytest.mark.parametrize("knn", KNN_MODES)
@pytest.mark.parametrize("N,D,n_ids,noise,k1,k2,seed", [
    (2048, 256, 64, 0.8, 30, 6, 0),
    (1500, 192, 1500, 1.0, 20, 6, 1),      # isotropic: worst case for top-k gaps
    (700, 64, 10, 0.3, 64, 6, 2),          # k1 at the 64-bit mask limit, tight clusters
    (300, 64, 10, 0.8, 7, 1, 3),           # k2 == 1, odd k1
])
def test_every_stage_against_oracle(N, D, n_ids, noise, k1, k2, seed, knn):
    import reid_gan_b200 as rg
    from oracle import rerank as orr
    x, _ = rg.synth(N, D, n_ids, noise, seed)
    st = rg.rerank_state(x.cuda(), k1, k2, knn=knn)
    o = orr.sparse_pipeline(x.numpy(), k1, k2)
    rank = st.rank.cpu().numpy()
    assert np.array_equal(rank, o["rank"]), "a1 neighbour lists"
    # a2: reciprocal sets as rank-position masks
    R = masks_to_lists(rank, st.R_mask.cpu().numpy())
    Rh = masks_to_lists(rank, st.Rh_mask.cpu().numpy())
    R_ref = orr.reciprocal_lists(o["rank"], k1)
    Rh_ref = orr.reciprocal_lists(o["rank"], orr.half_k(k1))
    for i in range(N):
        assert np.array_equal(R[i], R_ref[i]) and np.array_equal(Rh[i], Rh_ref[i]), "a2 row %d" % i
    # a3: expansion sets
    assert np.array_equal(st.E_ptr.cpu().numpy(), o["E_ptr"])
    nE = int(o["E_ptr"][-1])
    assert np.array_equal(st.E_idx.cpu().numpy()[:nE], o["E_idx"])
    # a4 / a5: weights
    np.testing.assert_allclose(st.V_val.cpu().numpy()[:nE], o["V_val"], atol=2e-6, rtol=0)
    assert np.array_equal(st.Q_ptr.cpu().numpy(), o["Vq_ptr"])
    nQ = int(o["Vq_ptr"][-1])
    assert np.array_equal(st.Q_idx.cpu().numpy()[:nQ], o["Vq_idx"])
    np.testing.assert_allclose(st.Q_val.cpu().numpy()[:nQ], o["Vq_val"], atol=2e-6, rtol=0)
    # a6: inverted index
    cp, ci, cv = orr.transpose_csr(st.Q_ptr.cpu().numpy(), st.Q_idx.cpu().numpy()[:nQ], st.Q_val.cpu().numpy()[:nQ], N)
    assert np.array_equal(st.C_ptr.cpu().numpy(), cp)
    assert np.array_equal(st.C_idx.cpu().numpy()[:nQ], ci)
    assert np.array_equal(st.C_val.cpu().numpy()[:nQ], cv)
    # a7 on the device's own V_qe: bit-exact against the oracle's sequential ascending-column sum
    jp, jj, jv = orr.jaccard_sparse(st.Q_ptr.cpu().numpy(), st.Q_idx.cpu().numpy()[:nQ],
                                    st.Q_val.cpu().numpy()[:nQ], N)
    Jd = rg.JaccardDistance(st).numpy()
    assert np.array_equal(Jd, orr.jaccard_dense_from_sparse(jp, jj, jv, N)), "a7 must be bit-exact given V_qe"


def test_stage_inputs_from_oracle_are_bit_exact():
    """Integer stages fed with the ORACLE's neighbour lists: outputs must be identical (no tolerance)."""
    import reid_gan_b200 as rg
    from reid_gan_b200 import _lib
    from reid_gan_b200._lib import call, ptr, stream_ptr
    from oracle import rerank as orr
    L = _lib.lib()
    N, D, k1 = 900, 64, 25
    x, _ = rg.synth(N, D, 30, 0.8, 11)
    rank = orr.exact_knn(x.numpy(), k1)
    d_rank = torch.from_numpy(rank.astype(np.int32)).cuda()
    for k in (k1, orr.half_k(k1)):
        m = torch.empty(N, dtype=torch.int64, device="cuda")
        call("reid_reciprocal_masks", ptr(d_rank), N, k1, k, 0, N, ptr(m), stream_ptr())
        ref = orr.k_reciprocal_masks(rank, k)
        got = m.cpu().numpy()
        for i in range(N):
            bits = [(int(got[i]) >> r) & 1 for r in range(ref.shape[1])]
            assert bits == list(ref[i].astype(int))
            assert (int(got[i]) & 0xFFFFFFFFFFFFFFFF) >> ref.shape[1] == 0


@pytest.mark.parametrize("eps", [0.3, 0.45, 0.6])
def test_dbscan_dense_matches_sklearn(eps):
    import reid_gan_b200 as rg
    from oracle import cluster as ocl
    rng = np.random.default_rng(int(eps * 100))
    pts = np.concatenate([rng.normal(c, 0.08, (60, 2)) for c in ((0, 0), (1, 0), (0, 1), (1, 1))] +
                         [rng.uniform(-0.5, 1.5, (80, 2))]).astype(np.float32)
    d = np.sqrt(((pts[:, None] - pts[None]) ** 2).sum(-1)).astype(np.float32)
    for ms in (2, 4, 9):
        ref = ocl.sklearn_dbscan(d, eps * 0.3, ms)
        c = rg.DBSCAN(eps=eps * 0.3, min_samples=ms)
        assert np.array_equal(c.fit_predict(d), ref)
        from sklearn.cluster import DBSCAN as SK
        sk = SK(eps=eps * 0.3, min_samples=ms, metric="precomputed").fit(d)
        assert np.array_equal(c.core_sample_indices_, sk.core_sample_indices_)
    # inclusive fp32 threshold: a distance of exactly float32(eps) is a neighbour
    d2 = np.full((5, 5), 2.0, np.float32)
    np.fill_diagonal(d2, 0.0)
    d2[0, 1] = d2[1, 0] = np.float32(0.6)
    assert np.array_equal(rg.DBSCAN(eps=0.6, min_samples=2).fit_predict(d2), ocl.sklearn_dbscan(d2, 0.6, 2))
    # all noise
    assert np.array_equal(rg.DBSCAN(eps=0.1, min_samples=3).fit_predict(d2), np.full(5, -1))


def test_full_size_properties():
    """N = 12,936 x 2048 (Market-1501 shape, BASELINE configs[0]): size-independent properties
    plus a sampled comparison with the oracle."""
    import reid_gan_b200 as rg
    from oracle import rerank as orr
    N, D, k1, k2 = 12936, 2048, 30, 6
    x, _ = rg.synth(N, D, 751, 0.8, 0)
    dist = rg.compute_jaccard_distance(x, k1=k1, k2=k2, print_flag=False)
    st = dist.state
    rank = st.rank.cpu().numpy()
    rows = np.arange(0, N, 97)
    assert np.array_equal(rank[rows], orr.exact_knn(x.numpy(), k1, rows=rows)), "neighbour lists (sampled rows)"
    qp = st.Q_ptr.cpu().numpy()
    qv = st.Q_val.cpu().numpy()[:qp[-1]]
    sums = np.add.reduceat(qv.astype(np.float64), qp[:-1])
    np.testing.assert_allclose(sums, 1.0, atol=1e-5)            # rows of V_qe sum to 1
    assert np.all(np.diff(qp) > 0)
    blk = dist.dense_device(0, 512).cpu().numpy()
    blk_t = torch.stack([dist.dense_device(i, i + 1)[0, :512] for i in range(0, 512, 64)]).cpu().numpy()
    assert np.array_equal(blk[0:512:64, :512], blk_t)             # row blocks are independent of blocking
    assert blk.min() >= 0.0 and blk.max() <= 1.0
    assert np.abs(np.diag(blk[:, :512])).max() <= 2e-6          # J_ii ~ 0
    sub = blk[:, :512]
    assert np.array_equal(sub, sub.T)                           # symmetric bit for bit
    labels = rg.DBSCAN(eps=0.6, min_samples=4, metric="precomputed", n_jobs=-1).fit_predict(dist)
    assert labels.shape == (N,) and labels.min() >= -1
    # labels from the sparse path == labels from the dense path on the same matrix block structure
    ncl = labels.max() + 1
    assert 300 < ncl < 1500
    # idempotence / determinism
    labels2 = rg.DBSCAN(eps=0.6, min_samples=4).fit_predict(rg.compute_jaccard_distance(x, k1, k2, print_flag=False))
    assert np.array_equal(labels, labels2)


# ---- symmetric tensor-core search (simgemm_sym.cu): bit-identical to the exact search --------------------
def _sym_case(N, D, n_ids, noise, seed, order=None, dup=0):
    import reid_gan_b200 as rg
    x, ids = rg.synth(N, D, n_ids, noise, seed)
    if dup:                                   # exact duplicates: equal keys, the index decides
        x[N - dup:] = x[:dup]
    if order == "sorted":                     # rows grouped by identity, like a dataset listed by person id
        x = x[torch.argsort(ids, stable=True)].contiguous()
    return x


@pytest.mark.parametrize("N,D,n_ids,noise,seed,order,dup,k", [
    (9000, 256, 9000, 1.0, 5, None, 0, 30),        # isotropic: smallest top-k gaps
    (8192, 128, 40, 0.5, 6, None, 0, 30),          # ~200 rows per identity: the k-th neighbour sits inside a dense cluster
    (10000, 192, 300, 0.8, 7, "sorted", 0, 20),    # identity-sorted rows: the threshold sample must not depend on row order
    (8500, 128, 280, 0.8, 8, None, 500, 30),       # 500 duplicated rows: ties broken by index
    (9000, 128, 100, 0.8, 12, None, 0, 64),        # k1 at the 64-bit mask limit: twice the candidates per row
    (8192, 256, 8192, 1.0, 13, None, 0, 48),       # isotropic, 32 < k <= 64
])
def test_symmetric_search_is_exact(monkeypatch, N, D, n_ids, noise, seed, order, dup, k):
    """Both layouts of the symmetric search: sample-first (the default: the prepass scores are handed to the main lists,
    the symmetric pass skips the sample blocks) and the separate-sample one (what the upload / sharded paths use)."""
    from reid_gan_b200 import faiss_rerank as fr, knn_tc
    x = _sym_case(N, D, n_ids, noise, seed, order, dup).cuda()
    ie, ke, _ = fr.knn_search(x, k, "exact")
    for sample_first in (True, False):
        monkeypatch.setattr(knn_tc, "SYM_SAMPLE_FIRST", sample_first)
        it, kt, info = fr.knn_search(x, k, "tc")
        assert info["mode"] == "tc-sym", "the symmetric kernel must be the one that ran"
        assert (info["sym"].get("layout") == "sample-first") == sample_first
        assert info["uncertified_rows"] == 0
        assert torch.equal(ie, it), "neighbour lists differ from the exact search"
        assert torch.equal(ke, kt), "keys differ from the exact search"
        cnt = info["cand_cnt"]
        assert int(cnt.min()) >= k, "every row must have kept at least k candidates"
        if sample_first:         # every pair is scored once: the lists are as long as in the other layout, tile count drops
            n_t, s_t = (N + 255) // 256, knn_tc.sample_size(N, k) // 256
            assert info["sym"]["tiles"] == (n_t - s_t) * (n_t - s_t + 1) // 2
            mean_sf = float(cnt.float().mean())
        else:
            assert abs(float(cnt.float().mean()) - mean_sf) <= 0.35 * mean_sf


@pytest.mark.parametrize("N,D,n_ids,k", [(8192 + 77, 128, 300, 30), (9000, 256, 100, 48)])
def test_symmetric_search_strip_flavour(monkeypatch, N, D, n_ids, k):
    """reid_knn_candidates_sym_wide (256 x 512 strips, both TMEM accumulators per unit; off by default -- measured
    slower): same neighbour lists as the exact search, and the pairing keeps every tile exactly once."""
    from reid_gan_b200 import faiss_rerank as fr, knn_tc
    t = knn_tc._tile_order(37, "cpu")
    u = knn_tc.pair_units(t)
    flat = [(int(a), int(b)) for a, b, c in u.tolist()] + [(int(a), int(c)) for a, b, c in u.tolist() if c >= 0]
    assert sorted(flat) == sorted(map(tuple, t.tolist()))
    x = _sym_case(N, D, n_ids, 0.8, 21).cuda()
    ie, ke, _ = fr.knn_search(x, k, "exact")
    from reid_gan_b200 import _lib
    monkeypatch.setattr(knn_tc, "SYM_WIDE", True)
    _lib.profiler.start()
    try:
        it, kt, info = fr.knn_search(x, k, "tc")
        torch.cuda.synchronize()
    finally:
        _lib.profiler.stop()
    assert info["mode"] == "tc-sym" and info["uncertified_rows"] == 0
    ran = _lib.profiler.summary()
    assert "reid_knn_candidates_sym_wide" in ran and "reid_knn_candidates_sym" not in ran, "the strip kernel must be the one that ran"
    assert torch.equal(ie, it) and torch.equal(ke, kt)


def test_symmetric_search_survives_bad_thresholds(monkeypatch):
    """A threshold sample that is far too optimistic (almost no column passes) must only cost time: the rows fail
    their certificate and are redone by the exact search."""
    from reid_gan_b200 import faiss_rerank as fr, knn_tc
    x = _sym_case(8192, 64, 260, 0.8, 9).cuda()
    ie, ke, _ = fr.knn_search(x, 20, "exact")
    monkeypatch.setattr(knn_tc, "SYM_TARGET", 16)          # ~16 candidates per row < k = 20
    it, kt, info = fr.knn_search(x, 20, "tc")
    assert info["mode"] == "tc-sym" and info["uncertified_rows"] > 0
    assert torch.equal(ie, it) and torch.equal(ke, kt)


def test_full_size_msmt17_shape_properties():
    """N = 32,621 x 2048 (BASELINE configs[1], the benchmark workload): properties that do not need the dense
    matrix, a sampled comparison of the neighbour lists with the oracle, and determinism of the whole pass."""
    import reid_gan_b200 as rg
    from reid_gan_b200 import pipeline
    from oracle import rerank as orr
    N, D, k1, k2 = 32621, 2048, 30, 6
    x, ids = rg.synth(N, D, 1041, 0.8, 0)
    xd = x.cuda()
    out = pipeline.pseudo_labels(xd, k1, k2, 0.6, 4, centroids=True)
    st = out["state"]
    assert st.knn_info["mode"] == "tc-sym" and st.knn_info["uncertified_rows"] == 0
    rank = st.rank.cpu().numpy()
    rows = np.arange(0, N, 409)
    assert np.array_equal(rank[rows], orr.exact_knn(x.numpy(), k1, rows=rows)), "neighbour lists (sampled rows)"
    assert np.array_equal(rank[:, 0], np.arange(N)), "a unit-norm row is its own nearest neighbour"
    qp = st.Q_ptr.cpu().numpy()
    qv = st.Q_val.cpu().numpy()[:qp[-1]]
    np.testing.assert_allclose(np.add.reduceat(qv.astype(np.float64), qp[:-1]), 1.0, atol=1e-5)
    qi = st.Q_idx.cpu().numpy()[:qp[-1]]
    assert all(np.all(np.diff(qi[qp[i]:qp[i + 1]]) > 0) for i in range(0, N, 997)), "V_qe rows sorted by column"
    # the eps-graph is symmetric (J is) and contains the diagonal
    labels = out["labels"].cpu().numpy()
    ncl = int(out["num_clusters"].item())
    assert ncl == labels.max() + 1
    # every pseudo-label is pure w.r.t. the generating identity on this well-separated set
    lab_ok = labels >= 0
    first = np.full(ncl, -1, np.int64)
    first[labels[lab_ok]] = ids.numpy()[lab_ok]
    assert np.array_equal(first[labels[lab_ok]], ids.numpy()[lab_ok])
    # cluster ids are numbered by ascending smallest member (sklearn's order)
    firsts = np.full(ncl, N, np.int64)
    np.minimum.at(firsts, labels[lab_ok], np.nonzero(lab_ok)[0])
    assert np.all(np.diff(firsts) > 0)
    cen = out["centroids"].cpu().numpy()
    np.testing.assert_allclose(np.linalg.norm(cen, axis=1), 1.0, atol=1e-5)
    out2 = pipeline.pseudo_labels(xd, k1, k2, 0.6, 4, centroids=True)
    assert torch.equal(out["labels"], out2["labels"]) and torch.equal(out["centroids"], out2["centroids"])
    assert torch.equal(st.rank, out2["state"].rank) and torch.equal(st.Q_val[:qp[-1]], out2["state"].Q_val[:qp[-1]])


@pytest.mark.parametrize("N,D,n_ids,k", [(3000, 128, 100, 15), (9000, 256, 300, 15), (1200, 64, 40, 80)])
def test_infomap_front_end(N, D, n_ids, k):
    """f1: get_dist_nbr / get_links (utils/infomap_cluster.py:230-234, 129-144) against the oracle: bit-exact."""
    import reid_gan_b200 as rg
    from reid_gan_b200 import infomap_cluster as ic
    from oracle import infomap as oi
    x, _ = rg.synth(N, D, n_ids, 0.8, 21)
    d, n = ic.get_dist_nbr(features=x.numpy(), k=k, knn_method='faiss-gpu')
    d_ref, n_ref = oi.get_dist_nbr(x.numpy(), k)
    assert d.dtype == np.float64 and n.dtype == np.int32
    assert np.array_equal(n, n_ref) and np.array_equal(d, d_ref)
    for min_sim in (0.3, 0.5):
        s, l = ic.get_links(single=[], links={}, nbrs=n, dists=d, min_sim=min_sim)
        s_ref, l_ref = oi.get_links(n_ref, d_ref, min_sim)
        assert s == s_ref and l == l_ref


@pytest.mark.parametrize("N,Q,D,n_ids,k1,k2,lam", [(900, 250, 64, 40, 20, 6, 0.3), (500, 120, 32, 30, 7, 1, 0.5),
                                                    (3000, 700, 128, 150, 20, 6, 0.3)])
def test_eval_rerank(N, Q, D, n_ids, k1, k2, lam):
    """f2: re_ranking (utils/rerank.py:31-97) against the oracle: neighbour lists and sets bit-exact, result within 1e-5."""
    import reid_gan_b200 as rg
    from oracle import eval_rerank as oe
    qg, qq, gg = oe.synthetic_distances(N, Q, D, n_ids, 31)
    ref, parts = oe.re_ranking(qg, qq, gg, k1, k2, lam, return_parts=True)
    # the comparison of the lists is only meaningful on tie-free fixtures (np.argsort is unstable at rerank.py:43)
    d = parts["dist"]
    srt = np.sort(d, axis=1)[:, :k1 + 2]
    assert (np.diff(srt, axis=1) > 0).all(), "fixture has ties inside the first k1+2 columns"
    got = rg.re_ranking(qg, qq, gg, k1=k1, k2=k2, lambda_value=lam)
    assert got.dtype == np.float32 and got.shape == (Q, N - Q)
    assert np.abs(got - ref).max() <= 1e-5
    # and through torch tensors, like evaluators.py would pass them without the .numpy()
    got2 = rg.re_ranking(torch.from_numpy(qg), torch.from_numpy(qq), torch.from_numpy(gg), k1=k1, k2=k2, lambda_value=lam)
    assert np.array_equal(got, got2)


@pytest.mark.parametrize("N,D,n_ids,k1", [(8192, 128, 260, 20), (10001, 192, 330, 30), (8200, 64, 8200, 12)])
def test_streamed_upload_matches_device_path(N, D, n_ids, k1):
    """Host-resident features (the reference's case, evaluators.py:19 `.cpu()`): the search that follows the chunked
    upload (knn_tc.knn_search_upload) must give the neighbour lists, V_qe and labels of the device-resident pass."""
    import reid_gan_b200 as rg
    from reid_gan_b200 import pipeline
    x, _ = rg.synth(N, D, n_ids, 0.8, 17)
    ref = pipeline.pseudo_labels(x.cuda(), k1, 6, 0.6, 4)
    for host in (x, x.pin_memory()):
        d = rg.compute_jaccard_distance(host, k1=k1, k2=6, print_flag=False)
        st = d.state
        assert st.knn_info["mode"] == "tc-sym" and st.knn_info["sym"].get("chunks")
        assert torch.equal(st.rank, ref["state"].rank)
        nq = ref["state"].q_total
        assert st.q_total == nq and torch.equal(st.Q_val[:nq], ref["state"].Q_val[:nq])
        labels = rg.DBSCAN(eps=0.6, min_samples=4).fit_predict(d)
        assert np.array_equal(labels, ref["labels"].cpu().numpy())


@pytest.mark.parametrize("N,Q,D,n_ids", [(2500, 600, 128, 80), (700, 150, 32, 30), (5000, 40, 64, 400)])
def test_evaluation_metrics(N, Q, D, n_ids):
    """f3: pairwise_distance / mean_ap / cmc against the oracle (metrics on the same matrix: 1e-12; distances 1e-4)."""
    import reid_gan_b200 as rg
    from reid_gan_b200 import evaluation as ev
    from oracle import ranking as orank
    x, ids = rg.synth(N, D, n_ids, 1.2, 5)
    cams = np.random.default_rng(5).integers(0, 4, N)
    q, g = x[:Q], x[Q:]
    qi, gi, qc, gc = ids[:Q].numpy(), ids[Q:].numpy(), cams[:Q], cams[Q:]
    feats = {"f%d" % i: x[i] for i in range(N)}
    query = [("f%d" % i, int(ids[i]), int(cams[i])) for i in range(Q)]
    gallery = [("f%d" % i, int(ids[i]), int(cams[i])) for i in range(Q, N)]
    dm, xq, yg = ev.pairwise_distance(feats, query, gallery)
    assert isinstance(dm, torch.Tensor) and dm.shape == (Q, N - Q) and np.array_equal(xq, q.numpy()) and np.array_equal(yg, g.numpy())
    d_ref = orank.pairwise_distance(q.numpy(), g.numpy())
    assert np.abs(dm.numpy() - d_ref).max() <= 1e-4
    d = dm.numpy().copy()
    d[:, 5] = d[:, 7]                                # force exact ties: same threshold for AP, index order for CMC
    assert abs(ev.mean_ap(d, qi, gi, qc, gc) - orank.mean_ap(d, qi, gi, qc, gc)) <= 1e-12
    for kw in (dict(first_match_break=True), dict(), dict(separate_camera_set=True, first_match_break=True)):
        a = ev.cmc(d, qi, gi, qc, gc, topk=50, **kw)
        b = orank.cmc(d, qi, gi, qc, gc, topk=50, **kw)
        assert np.abs(a - b).max() <= 1e-12
    assert abs(ev.mean_ap(d[:, :Q]) - orank.mean_ap(d[:, :Q])) <= 1e-12          # default ids / cameras


def test_cmc_single_gallery_shot_follows_the_reference_rng_protocol():
    """ranking.py:53-66 draws one gallery instance per identity with np.random.choice, 10 times per query: with the same
    seed the drop-in consumes the generator in the same order and returns the reference's numbers."""
    from reid_gan_b200 import evaluation as ev
    g = np.load(os.path.join(GOLD, "ranking_q120_g380.npz"))
    s = np.load(os.path.join(GOLD, "ranking_sgs_q120_g380.npz"))
    args = (g["distmat"], g["q_ids"], g["g_ids"], g["q_cams"], g["g_cams"])
    for name, kw in (("sgs", dict()), ("sgs_fmb", dict(first_match_break=True)), ("sgs_sep", dict(separate_camera_set=True))):
        np.random.seed(1234)
        got = ev.cmc(*args, topk=50, single_gallery_shot=True, **kw)
        assert np.abs(got - s["cmc_" + name + "_seed1234"]).max() <= 1e-12


@pytest.mark.parametrize("ci,N,D,n_ids,noise", [
    (0, 8193, 64, 264, 0.8), (1, 8447, 192, 8447, 1.0), (2, 9000, 192, 8, 0.3), (3, 12345, 64, 398, 0.8),
    (4, 16384, 64, 8, 0.3), (5, 20001, 192, 20001, 1.0), (6, 8192, 192, 8, 0.3), (7, 20001, 64, 645, 0.8)])
def test_symmetric_search_random_shapes(ci, N, D, n_ids, noise):
    """Shapes that are not multiples of the 128 / 256-row tiles, huge near-duplicate clusters (windows that overflow
    and fall back to the exact search), duplicated rows, rows that are not unit norm, k from 1 to 32."""
    import reid_gan_b200 as rg
    from reid_gan_b200 import faiss_rerank as fr
    k = (1, 5, 30, 32)[ci % 4]
    x, _ = rg.synth(N, D, n_ids, noise, ci)
    if ci % 3 == 0:
        x[N // 2: N // 2 + 300] = x[:300]
    if ci % 5 == 0:
        g = torch.Generator().manual_seed(ci)
        x = x * (0.5 + torch.rand(N, 1, generator=g))
    xd = x.cuda()
    ie, ke, _ = fr.knn_search(xd, k, "exact", metric="ip")        # the inner-product key on both sides (the tensor-core
    it, kt, info = fr.knn_search(xd, k, "tc", metric="ip")         # kernel under test), whatever the norms
    assert info["mode"] == "tc-sym"
    assert torch.equal(ie, it) and torch.equal(ke, kt)


@pytest.mark.parametrize("N,D,k,mode", [(3000, 128, 20, "auto"), (9000, 64, 30, "auto"), (2500, 96, 10, "exact")])
def test_rows_of_different_norms_are_ranked_by_l2_like_the_reference(N, D, k, mode):
    """faiss IndexFlatL2 (faiss_rerank.py:58-62) ranks by squared L2.  For rows that do not share one norm that is NOT the
    inner-product order: the search must notice (device-side min / max of the squared norms) and rank by the L2 key;
    the whole pass then matches the oracle, which takes the same decision (oracle.rerank.exact_knn metric='auto')."""
    import reid_gan_b200 as rg
    from reid_gan_b200 import faiss_rerank as fr, pipeline
    from oracle import rerank as orr, cluster as ocl
    x, _ = rg.synth(N, D, max(1, N // 25), 0.8, 11)
    g = torch.Generator().manual_seed(5)
    x = (x * (0.6 + 0.8 * torch.rand(N, 1, generator=g))).contiguous()
    ref_l2, ref_key = orr.exact_knn(x.numpy(), k, return_keys=True, metric="l2")
    ref_ip = orr.exact_knn(x.numpy(), k, metric="ip")
    assert not np.array_equal(ref_l2, ref_ip), "the fixture must tell the two orders apart"
    idx, key, info = fr.knn_search(x.cuda(), k, mode)
    assert info["metric"] == "l2" and info["mode"] == "exact-l2"
    assert np.array_equal(idx.cpu().numpy(), ref_l2)
    assert np.array_equal(key.cpu().numpy(), ref_key)
    idx_ip, _, info_ip = fr.knn_search(x.cuda(), k, mode, metric="ip")        # what get_dist_nbr (IndexFlatIP) asks for
    assert info_ip["metric"] == "ip" and np.array_equal(idx_ip.cpu().numpy(), ref_ip)
    # the whole pass: speculative single-GPU flavour (finish() notices the norms), the drop-in call, and the oracle
    out = pipeline.pseudo_labels(x.cuda(), k, 6 if k >= 6 else 1, 0.6, 4)
    assert out["state"].knn_info["metric"] == "l2"
    assert np.array_equal(out["state"].rank.cpu().numpy(), ref_l2)
    J_ref = orr.compute_jaccard_distance_oracle(x.numpy(), k, 6 if k >= 6 else 1, rank=ref_l2)
    J = np.asarray(rg.compute_jaccard_distance(x, k1=k, k2=6 if k >= 6 else 1, print_flag=False))
    assert np.abs(J - J_ref).max() <= 1e-5


# ---- CUDA path against the golden vectors of the "next" rows (outputs of the unmodified reference) -----------
def test_golden_infomap_front_end_cuda():
    from reid_gan_b200 import infomap_cluster as ic
    g = np.load(os.path.join(GOLD, "infomap_n400_k15.npz"))
    d, n = ic.get_dist_nbr(features=g["x"], k=int(g["k"]), knn_method='faiss-gpu')
    assert d.dtype == g["dists"].dtype and n.dtype == g["nbrs"].dtype
    assert np.array_equal(n, g["nbrs"]) and np.array_equal(d, g["dists"])
    for min_sim, tag in ((0.3, "ms30"), (0.5, "ms50")):
        single, links = ic.get_links(single=[], links={}, nbrs=n, dists=d, min_sim=min_sim)
        ref = {(int(a), int(b)): float(w) for (a, b), w in zip(g["links_ij_" + tag], g["links_w_" + tag])}
        assert links == ref and single == [int(v) for v in g["single_" + tag]]


def test_golden_eval_rerank_cuda():
    import reid_gan_b200 as rg
    g = np.load(os.path.join(GOLD, "evalrerank_q100_g260.npz"))
    for k1, k2 in ((20, 6), (7, 1)):
        ref = g["final_k%d_%d" % (k1, k2)]
        got = rg.re_ranking(g["q_g"], g["q_q"], g["g_g"], k1=k1, k2=k2, lambda_value=float(g["lambda_k%d_%d" % (k1, k2)]))
        assert got.dtype == ref.dtype and got.shape == ref.shape and np.abs(got - ref).max() <= 1e-5


def test_golden_ranking_metrics_cuda():
    from reid_gan_b200 import evaluation as ev
    g = np.load(os.path.join(GOLD, "ranking_q120_g380.npz"))
    d = ev.pairwise_distance_device(torch.from_numpy(g["q"]).cuda(), torch.from_numpy(g["g"]).cuda()).cpu().numpy()
    assert np.abs(d - g["distmat"]).max() <= 1e-4
    args = (g["distmat"], g["q_ids"], g["g_ids"], g["q_cams"], g["g_cams"])
    assert abs(ev.mean_ap(*args) - float(g["mAP"])) <= 1e-12
    assert np.abs(ev.cmc(*args, topk=50, first_match_break=True) - g["cmc_market"]).max() <= 1e-12
    assert np.abs(ev.cmc(*args, topk=50) - g["cmc_allshots"]).max() <= 1e-12
    assert np.abs(ev.cmc(*args, topk=50, separate_camera_set=True, first_match_break=True) - g["cmc_sepcam"]).max() <= 1e-12


HALF = sorted(glob.glob(os.path.join(GOLD, "halfrerank_*.npz")))


@pytest.mark.parametrize("path", HALF, ids=[os.path.basename(p)[:-4] for p in HALF])
def test_use_float16_against_reference(path):
    """compute_jaccard_distance(..., use_float16=True) (faiss_rerank.py:37): float16 matrix, the reference's float16
    rounding points.  Bars as in tests/test_oracle.py: identical sparsity, >= 99.9 % of the entries identical to the
    unmodified reference's, the rest within a few float16 ulps; labels equal sklearn's on the reference's matrix."""
    import reid_gan_b200 as rg
    from oracle import cluster as ocl
    g = np.load(path)
    x, k1, k2 = torch.from_numpy(g["x"]), int(g["k1"]), int(g["k2"])
    N = x.shape[0]
    J_ref = np.ones((N, N), dtype=np.float16)
    J_ref[g["J_rows"], g["J_cols"]] = g["J_vals"]
    d = rg.compute_jaccard_distance(x, k1=k1, k2=k2, print_flag=False, search_option=3, use_float16=True)
    assert d.dtype == np.float16
    J = np.asarray(d)
    assert J.dtype == np.float16 and J.shape == (N, N)
    assert np.array_equal(J == 1.0, J_ref == 1.0)
    assert (J != J_ref).mean() <= 1e-3
    assert np.abs(J.astype(np.float32) - J_ref.astype(np.float32)).max() <= 2e-3
    assert np.array_equal(J, J.T)
    # the sparse DBSCAN path applies the same float16 roundings as the dense matrix it never builds
    lab_sparse = rg.DBSCAN(eps=0.6, min_samples=4, metric="precomputed").fit_predict(d)
    lab_dense = ocl.sklearn_dbscan(J.astype(np.float64), 0.6, 4)
    assert np.array_equal(lab_sparse, lab_dense)


def test_device_resident_feature_hand_off():
    """f4 (clustercontrast/evaluators.py:16-68 + train_usl.py:152-153): the same extract_features call, but the rows never
    visit the host; features.matrix(sorted names) equals the reference flow's host-side torch.cat bit for bit, and
    the pass started from it gives the same labels as the pass started from the host matrix."""
    import reid_gan_b200 as rg
    from reid_gan_b200 import pipeline
    torch.manual_seed(0)
    N, Din, D, B = 9000, 48, 256, 512
    model = torch.nn.Sequential(torch.nn.Linear(Din, D)).cuda()

    class Net(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.body = model

        def forward(self, x):
            return torch.nn.functional.normalize(self.body(x), dim=1)

    net = Net().cuda()
    centres = torch.randn(300, Din)
    ids = torch.randint(0, 300, (N,))
    imgs = centres[ids] + 0.35 * torch.randn(N, Din)
    names = ["img_%05d.jpg" % i for i in torch.randperm(N).tolist()]           # loader order != sorted order
    loader = [(imgs[a:a + B], names[a:a + B], ids[a:a + B].tolist(), None, None) for a in range(0, N, B)]
    feats, labels = rg.extract_features(net, loader, print_freq=10 ** 9)
    assert len(feats) == N and list(labels.keys()) == names and feats[names[3]].is_cuda
    order = sorted(names)
    x_dev = feats.matrix(order)
    # the reference flow: per-batch .cpu(), dict of rows, N-way cat on the host
    ref = {}
    with torch.no_grad():
        for im, fn, _, _, _ in loader:
            out = net(im.cuda()).data.cpu()
            for f, o in zip(fn, out):
                ref[f] = o
    x_ref = torch.cat([ref[f].unsqueeze(0) for f in order], 0)
    assert x_dev.is_cuda and torch.equal(x_dev.cpu(), x_ref)
    a = pipeline.pseudo_labels(x_dev, 20, 6, 0.6, 4)["labels"].cpu().numpy()
    b = rg.DBSCAN(eps=0.6, min_samples=4, metric="precomputed").fit_predict(
        rg.compute_jaccard_distance(x_ref, k1=20, k2=6, print_flag=False))
    assert np.array_equal(a, b) and a.max() > 10


def test_scan_counts_is_64_bit_and_multi_tile():
    """reid_scan_counts: decoupled look-back over many CTAs, sums in 64 bits -- counts that add up to more than 2^31 inside
    one 2048-element tile (the slot bounds of reid_jaccard_bounds can) must not wrap; ragged sizes; repeated use of the
    self-resetting state."""
    from reid_gan_b200.faiss_rerank import _scan_async
    g = torch.Generator().manual_seed(0)
    for n, hi in ((1, 5), (2047, 100), (2048, 100), (2049, 100), (70001, 3_000_000), (300000, 40)):
        cnt = torch.randint(0, hi, (n,), generator=g, dtype=torch.int32)
        if n == 70001:
            cnt[:2048] = 2_000_000                      # 4.1e9 inside the first tile alone
        ptr, stats = _scan_async(cnt.cuda(), n, torch.device("cuda"))
        ref = torch.zeros(n + 1, dtype=torch.int64)
        ref[1:] = torch.cumsum(cnt.to(torch.int64), 0)
        assert torch.equal(ptr.cpu(), ref)
        assert stats.tolist() == [int(ref[-1]), int(cnt.max()), int((cnt.to(torch.int64) ** 2).sum())]


def test_centroids_of_very_large_clusters():
    """train_usl.py:169-191 with clusters far beyond the shared-memory member list (the kernel then streams the labels):
    same means as the oracle, outliers (-1) skipped, empty label ids give zero rows."""
    import reid_gan_b200 as rg
    from oracle import cluster as ocl
    g = torch.Generator().manual_seed(3)
    x = torch.randn(15000, 192, generator=g)
    labels = torch.randint(-1, 3, (15000,), generator=g).numpy()        # three clusters of ~3750 rows + outliers
    cen = rg.generate_cluster_features(labels, x, normalize=True).cpu().numpy()
    np.testing.assert_allclose(cen, ocl.cluster_centroids(x.numpy(), labels), rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize("splits", [2, 4])
def test_prepass_thresholds_with_column_splits(splits):
    """The sampling prepass of a row block whose 256-row units do not fill the CTA-pair slots is split over the sample
    columns (sharded.knn_search_tiles at W >= 4).  tau_i = r-th best sample score over all of a row's lists: never above
    the exact r-th best (a threshold can only err towards more candidates) and equal to it (to the 16 resolved bits) on
    practically every row, for the unsplit and the split prepass alike."""
    import ctypes
    import reid_gan_b200 as rg
    from reid_gan_b200 import knn_tc as kt
    from reid_gan_b200._lib import call, ptr, stream_ptr
    N, D, k = 8192, 512, 30
    x = rg.synth(N, D, 300, 0.8, 1)[0].cuda()
    dev, sp = x.device, stream_ptr()
    xh = torch.empty((N, D), dtype=torch.float16, device=dev)
    call("reid_features_to_half", ptr(x), N, D, kt.SCALE_LOG2, ptr(xh), None, sp)
    m = kt.sample_size(N, k)
    xs = torch.empty((m, D), dtype=torch.float16, device=dev)
    call("reid_features_sample", ptr(xh), N, D, m, kt._sample_stride(N, m), ptr(xs), sp)
    r = kt.sym_rank(k)
    b0, b1 = 1024, 1024 + 2560 + 77                      # a ragged block in the middle
    nb = b1 - b0
    assert kt.prepass_splits(4096, 2048) == 4 and kt.prepass_splits(8192, 2048) == 2 and kt.prepass_splits(16384, 2048) == 1
    # exact r-th best sample score of every row, from the same fp16 operands (fp32 accumulation like the tensor core)
    s = (xh[b0:b1].float() @ xs.float().T) * 2.0 ** (-2 * kt.SCALE_LOG2)
    exact = torch.sort(s, dim=1, descending=True).values[:, r - 1]
    pre = torch.empty(nb * 2 * splits * kt.TC_CAP, dtype=torch.int64, device=dev)
    pre_cnt = torch.zeros(nb * 2 * splits, dtype=torch.int32, device=dev)
    pre_tau = torch.empty(nb, dtype=torch.int32, device=dev)
    call("reid_knn_candidates_tc_ab", ptr(xh), N, ptr(xs), m, D, kt.SCALE_LOG2, b0, b1, -r, splits, 2, ptr(pre), ptr(pre_cnt),
         ptr(pre_tau), sp)
    tau = torch.empty(nb, dtype=torch.float32, device=dev)
    tau_ord = torch.empty(nb, dtype=torch.int32, device=dev)
    call("reid_knn_sample_tau", ptr(pre), ptr(pre_cnt), ptr(pre_tau), 2 * splits, nb, r, ptr(tau), ptr(tau_ord), sp)
    torch.cuda.synchronize()
    assert bool(torch.isfinite(tau).all())
    assert bool((tau <= exact + 1e-6).all())
    close = (exact - tau) <= 2.0 ** -7 * exact.abs().clamp_min(1e-3)      # 16 resolved bits of the fp32 image = 8 mantissa bits
    assert float(close.float().mean()) > 0.99
