"""ORACLE (test infrastructure, never on the product path).

CPU restatement of the evaluation metrics:
clustercontrast/evaluators.py pairwise_distance :71-88 and clustercontrast/evaluation_metrics/ranking.py
mean_ap :82-115, cmc :18-79 (single_gallery_shot, which draws from np.random, is not covered).

Instead of sorting every row (np.argsort at ranking.py:41,104 is unstable, so ties are unordered in the reference) the
restatement uses what the metrics actually depend on: for every positive (valid match) gallery item p of a query,
    le(p)  = number of valid gallery items with distance <= d_p        (precision denominator, ties included)
    tp(p)  = number of valid positives with distance <= d_p
    pos(p) = number of valid items strictly before p in (distance, index) order   (its 0-based rank for CMC)
average_precision_score (sklearn, distinct thresholds) is then  sum_p tp(p) / le(p) / P,  and CMC reads pos(p).
Pinned in tests/test_oracle.py against the reference functions run verbatim (oracle/ref_shim.load_ranking()).
"""
import numpy as np


def pairwise_distance(x, y):
    """evaluators.py:81-87: ||x||^2 + ||y||^2 - 2 x.y^T, float32."""
    x = np.asarray(x, dtype=np.float32)
    y = np.asarray(y, dtype=np.float32)
    xx = (x * x).sum(axis=1, dtype=np.float32)[:, None]
    yy = (y * y).sum(axis=1, dtype=np.float32)[None, :]
    return (xx + yy - np.float32(2.0) * (x @ y.T)).astype(np.float32)


def _defaults(m, n, query_ids, gallery_ids, query_cams, gallery_cams):
    if query_ids is None:
        query_ids = np.arange(m)
    if gallery_ids is None:
        gallery_ids = np.arange(n)
    if query_cams is None:
        query_cams = np.zeros(m).astype(np.int32)
    if gallery_cams is None:
        gallery_cams = np.ones(n).astype(np.int32)
    return np.asarray(query_ids), np.asarray(gallery_ids), np.asarray(query_cams), np.asarray(gallery_cams)


def _positive_stats(d, valid, match):
    """For one query row: (le, tp, pos) of every valid positive, see the module docstring."""
    pidx = np.nonzero(valid & match)[0]
    dv = d[valid]
    iv = np.nonzero(valid)[0]
    le = np.array([(dv <= d[p]).sum() for p in pidx], dtype=np.int64)
    tp = np.array([(d[pidx] <= d[p]).sum() for p in pidx], dtype=np.int64)
    pos = np.array([((dv < d[p]) | ((dv == d[p]) & (iv < p))).sum() for p in pidx], dtype=np.int64)
    return pidx, le, tp, pos


def mean_ap(distmat, query_ids=None, gallery_ids=None, query_cams=None, gallery_cams=None):
    distmat = np.asarray(distmat)
    m, n = distmat.shape
    query_ids, gallery_ids, query_cams, gallery_cams = _defaults(m, n, query_ids, gallery_ids, query_cams, gallery_cams)
    aps = []
    for i in range(m):
        valid = (gallery_ids != query_ids[i]) | (gallery_cams != query_cams[i])      # :108-109
        match = gallery_ids == query_ids[i]                                          # :105
        pidx, le, tp, _ = _positive_stats(distmat[i], valid, match)
        if pidx.size == 0:                                                           # :112
            continue
        aps.append(float(np.sum(tp.astype(np.float64) / le.astype(np.float64)) / pidx.size))
    if len(aps) == 0:
        raise RuntimeError("No valid query")
    return np.mean(aps)


def cmc(distmat, query_ids=None, gallery_ids=None, query_cams=None, gallery_cams=None, topk=100,
        separate_camera_set=False, single_gallery_shot=False, first_match_break=False):
    if single_gallery_shot:
        raise NotImplementedError("single_gallery_shot samples with np.random (ranking.py:10-16): not reproducible")
    distmat = np.asarray(distmat)
    m, n = distmat.shape
    query_ids, gallery_ids, query_cams, gallery_cams = _defaults(m, n, query_ids, gallery_ids, query_cams, gallery_cams)
    ret = np.zeros(topk)
    num_valid_queries = 0
    for i in range(m):
        valid = (gallery_ids != query_ids[i]) | (gallery_cams != query_cams[i])      # :47-48
        if separate_camera_set:
            valid &= gallery_cams != query_cams[i]                                   # :49-51
        match = gallery_ids == query_ids[i]
        pidx, _, _, pos = _positive_stats(distmat[i], valid, match)
        if pidx.size == 0:                                                           # :52
            continue
        index = np.sort(pos)                                                         # positions of the matches among the valid items
        delta = 1. / len(index)
        for j, k in enumerate(index):                                                # :70-75
            if k - j >= topk:
                break
            if first_match_break:
                ret[k - j] += 1
                break
            ret[k - j] += delta
        num_valid_queries += 1
    if num_valid_queries == 0:
        raise RuntimeError("No valid query")
    return ret.cumsum() / num_valid_queries
