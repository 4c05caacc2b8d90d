"""ORACLE (test infrastructure, never on the product path).

CPU restatement of the evaluation-time k-reciprocal re-ranking:
clustercontrast/utils/rerank.py re_ranking :31-97, called by Evaluator.evaluate(rerank=True)
(clustercontrast/evaluators.py:138-142).  Same k-reciprocal / Jaccard mathematics as the pseudo-label path, but on
a (query+gallery)^2 matrix of column-max-normalised squared distances, with exp(-d) weights and the final blend
(1 - lambda) * Jaccard + lambda * original.  The set logic and the sparse stages reuse oracle/rerank.py.

np.argsort at :42 is numpy's default (unstable) sort: ties inside the first k1+1 columns of a row are not ordered
by the reference; the restatement (and the CUDA path) order them by index, and fixtures are required to be tie free
there.  Pinned in tests/test_oracle.py against the reference function run verbatim (oracle/ref_shim.load_eval_rerank()).
"""
import numpy as np

from . import rerank as orr


def normalised_distance(q_g_dist, q_q_dist, g_g_dist):
    """rerank.py:36-41 -> float32 (M, M), M = Q + G: row i = column i of the squared block matrix / its maximum."""
    original_dist = np.concatenate(
        [np.concatenate([q_q_dist, q_g_dist], axis=1),
         np.concatenate([q_g_dist.T, g_g_dist], axis=1)], axis=0)
    original_dist = np.power(original_dist, 2).astype(np.float32)
    return np.transpose(1. * original_dist / np.max(original_dist, axis=0))


def initial_rank(dist, cols):
    """rerank.py:43, first `cols` columns: ascending distance, ties by index."""
    M = dist.shape[0]
    out = np.empty((M, cols), dtype=np.int64)
    for i in range(M):
        part = np.argpartition(dist[i], cols - 1)[:cols]
        kth = dist[i, part].max()
        cand = np.nonzero(dist[i] <= kth)[0]
        o = np.lexsort((cand, dist[i, cand]))[:cols]
        out[i] = cand[o]
    return out


def re_ranking(q_g_dist, q_q_dist, g_g_dist, k1=20, k2=6, lambda_value=0.3, return_parts=False):
    q_g_dist = np.asarray(q_g_dist)
    Q = q_g_dist.shape[0]
    dist = np.ascontiguousarray(normalised_distance(q_g_dist, np.asarray(q_q_dist), np.asarray(g_g_dist)))
    M = dist.shape[0]
    rank = initial_rank(dist, k1 + 1)                                        # :43, only [:k1+1] / [:k2] are read
    ep, ei = orr.expand(rank, k1)                                            # :50-64 (slices :k1+1 and :around(k1/2)+1)
    rows = np.repeat(np.arange(M), np.diff(ep))
    w = np.exp(-dist[rows, ei]).astype(np.float32)                           # :66
    sums = np.add.reduceat(w.astype(np.float64), ep[:-1]).astype(np.float32)
    ev = (w / sums[rows]).astype(np.float32)                                 # :67 (sum order: see the tolerance note)
    if k2 != 1:
        qp, qi, qv = orr.query_expand(ep, ei, ev, rank, k2)                  # :69-74
    else:
        qp, qi, qv = ep, ei, ev
    jp, jj, jv = orr.jaccard_sparse(qp, qi, qv, M, row_begin=0, row_end=Q)   # :75-93
    J = orr.jaccard_dense_from_sparse(jp, jj, jv, M, row_begin=0)[:Q]
    final = J * np.float32(1 - lambda_value) + dist[:Q] * np.float32(lambda_value)   # :95
    final = final[:, Q:].astype(np.float32)
    if return_parts:
        return final, dict(dist=dist, rank=rank, E=(ep, ei), V=ev, Q=(qp, qi, qv))
    return final


def synthetic_distances(N, Q, D, n_ids, seed):
    """Test helper: (q_g, q_q, g_g) squared-L2 matrices of synthetic features, computed the way
    clustercontrast/evaluators.py:78-87 pairwise_distance does (expand + addmm_)."""
    import torch
    from reid_gan_b200.synth import synth
    x, _ = synth(N, D, n_ids, 0.8, seed)

    def pdist(a, b):
        m, n = a.size(0), b.size(0)
        d = torch.pow(a, 2).sum(dim=1, keepdim=True).expand(m, n) + torch.pow(b, 2).sum(dim=1, keepdim=True).expand(n, m).t()
        d = d.clone()
        d.addmm_(a, b.t(), beta=1, alpha=-2)
        return d
    q, g = x[:Q], x[Q:]
    return pdist(q, g).numpy(), pdist(q, q).numpy(), pdist(g, g).numpy()
