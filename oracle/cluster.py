"""ORACLE (test infrastructure, never on the product path).

CPU restatement of the clustering half of the pseudo-label pass:
  * DBSCAN(eps, min_samples, metric='precomputed').fit_predict
    -- call site examples/cluster_contrast_train_usl.py:160,163; algorithm in
       scikit-learn (third party, un-pinned in reference setup.py:13; 1.9.0 in this
       image): sklearn/cluster/_dbscan.py:397-475 + _dbscan_inner.pyx.
  * generate_cluster_features + F.normalize
    -- examples/cluster_contrast_train_usl.py:169-182, 191.

Pinned in tests/test_oracle.py against sklearn.cluster.DBSCAN itself (which IS the
reference implementation, importable here and on the GPU box) and against the
reference closure run through oracle/ref_shim.py.
"""
import numpy as np


def dbscan_from_neighbors(nbr_ptr, nbr_idx, min_samples):
    """Labels from eps-neighbourhood lists (self included, as sklearn's
    radius_neighbors gives for precomputed input).

    sklearn semantics restated order-free: core <=> |nbr| >= min_samples; clusters =
    connected components of the core-core graph, numbered 0,1,.. by ascending smallest
    member index (dbscan_inner scans i ascending and opens a new label at every
    unlabelled core); a non-core point takes the smallest label among adjacent cores
    (first DFS to reach it wins; clusters are grown in label order); else -1."""
    import scipy.sparse as sp
    from scipy.sparse.csgraph import connected_components
    N = nbr_ptr.size - 1
    cnt = np.diff(nbr_ptr)
    core = cnt >= min_samples
    rows = np.repeat(np.arange(N), cnt)
    cc = core[rows] & core[nbr_idx]
    G = sp.csr_matrix((np.ones(int(cc.sum()), np.int8), (rows[cc], nbr_idx[cc])), shape=(N, N))
    _, comp = connected_components(G, directed=False)
    labels = np.full(N, -1, dtype=np.int64)
    core_idx = np.nonzero(core)[0]
    if core_idx.size:
        # component id -> rank of its smallest core index
        first = {}
        for i in core_idx:
            first.setdefault(comp[i], len(first))
        labels[core_idx] = [first[comp[i]] for i in core_idx]
    # border points
    nb_edge = (~core[rows]) & core[nbr_idx]
    if nb_edge.any():
        br, bl = rows[nb_edge], labels[nbr_idx[nb_edge]]
        best = np.full(N, np.iinfo(np.int64).max)
        np.minimum.at(best, br, bl)
        hit = best != np.iinfo(np.int64).max
        labels[hit & ~core] = best[hit & ~core]
    return labels, core


def dbscan_dense(dist, eps, min_samples):
    """Restatement on a dense fp32 matrix: neighbourhood = d <= eps compared in the
    matrix dtype, inclusive, self included (sklearn/neighbors/_base.py radius_neighbors
    on precomputed input)."""
    d = np.asarray(dist)
    thr = d.dtype.type(eps)
    N = d.shape[0]
    ptr = [0]
    idx = []
    for i in range(N):
        nb = np.nonzero(d[i] <= thr)[0]
        idx.append(nb)
        ptr.append(ptr[-1] + nb.size)
    return dbscan_from_neighbors(np.asarray(ptr, np.int64),
                                 np.concatenate(idx) if idx else np.zeros(0, np.int64), min_samples)[0]


def dbscan_sparse_J(jp, jj, jv, eps, min_samples):
    """From the sparse J of oracle/rerank.jaccard_sparse (pairs without a shared column
    have J == 1 and are never neighbours for eps < 1)."""
    assert eps < 1.0
    keep = jv <= np.float32(eps)
    N = jp.size - 1
    rows = np.repeat(np.arange(N), np.diff(jp))
    cnt = np.bincount(rows[keep], minlength=N)
    ptr = np.concatenate(([0], np.cumsum(cnt))).astype(np.int64)
    return dbscan_from_neighbors(ptr, jj[keep], min_samples)[0]


def sklearn_dbscan(dist, eps, min_samples):
    """The reference implementation itself (train_usl.py:160,163)."""
    from sklearn.cluster import DBSCAN
    return DBSCAN(eps=eps, min_samples=min_samples, metric="precomputed", n_jobs=-1).fit_predict(dist)


def cluster_centroids(x, labels):
    """train_usl.py:169-182 + F.normalize (:191): per-label mean in ascending label
    order (-1 skipped), then L2 normalise with eps 1e-12.  fp32."""
    x = np.asarray(x, dtype=np.float32)
    labels = np.asarray(labels)
    labs = np.unique(labels[labels >= 0])
    out = np.empty((labs.size, x.shape[1]), dtype=np.float32)
    for k, lab in enumerate(labs):
        members = x[labels == lab]
        acc = np.zeros(x.shape[1], dtype=np.float32)
        for m in members:                                      # sequential fp32 adds
            acc = acc + m
        mu = acc / np.float32(members.shape[0])
        nrm = np.sqrt(np.sum(mu.astype(np.float32) ** 2, dtype=np.float32))
        out[k] = mu / max(nrm, np.float32(1e-12))
    return out
