"""ORACLE (test infrastructure, never on the product path).

CPU restatement of the kNN front end of the Infomap clustering variant:
clustercontrast/utils/infomap_cluster.py -- knn_faiss :51-78 (IndexFlatIP self-search, distances 1 - sim),
knns2ordered_nbrs :114-125, get_dist_nbr :230-234, get_links :129-144.  Infomap itself (the `infomap` package) is
third-party and out of scope.

faiss is un-vendored and un-pinned (CC/setup.py:13); its contract at the call site is "k largest inner products,
descending".  As for the L2 search of the DBSCAN variant, the order is anchored on the canonical key
fp32(sum_d x_id * x_jd accumulated in fp64), ties by smaller index.

Pinned in tests/test_oracle.py against the reference functions run verbatim through oracle/ref_shim.load_infomap_cluster().
"""
import numpy as np

from .rerank import exact_knn


def get_dist_nbr(features, k=80):
    """infomap_cluster.py:230-234 -> (dists float32 (N, k) ascending, nbrs int32 (N, k))."""
    x = np.ascontiguousarray(np.asarray(features), dtype=np.float32)           # :59 feats.astype('float32')
    nbrs, sims = exact_knn(x, k, return_keys=True)                             # :73 index.search(feats, k)
    dists = (np.float32(1.0) - sims.astype(np.float32)).astype(np.float32)     # :75 1 - np.array(sim, float32)
    order = np.argsort(dists, axis=1, kind="stable")                           # :120-123 (already ascending: a no-op)
    rows = np.arange(x.shape[0])[:, None]
    # np.array(knns) at :116 mixes the int32 and float32 halves of every tuple, so the reference hands out float64
    return dists[rows, order].astype(np.float64), nbrs.astype(np.int32)[rows, order]


def get_links(nbrs, dists, min_sim):
    """infomap_cluster.py:129-144 -> (single: list of rows without a link, links: {(i, j): similarity})."""
    single, links = [], {}
    for i in range(nbrs.shape[0]):
        count = 0
        for j in range(nbrs.shape[1]):
            if i == nbrs[i][j]:
                pass
            elif dists[i][j] <= 1 - min_sim:
                count += 1
                links[(i, int(nbrs[i][j]))] = float(1 - dists[i][j])
            else:
                break
        if count == 0:
            single.append(i)
    return single, links
