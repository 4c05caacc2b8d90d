"""The whole pseudo-label pass of one epoch as a single device-resident call:

    rerank_dist   = compute_jaccard_distance(features, k1, k2)          # train_usl.py:154
    pseudo_labels = DBSCAN(eps, min_samples=4, 'precomputed').fit_predict(rerank_dist)   # :160-163
    centers       = F.normalize(generate_cluster_features(pseudo_labels, features))       # :169-191

`pseudo_labels(...)` runs a1-a9 back to back on the current device without ever forming the
N x N matrix; it is what `compute_jaccard_distance` + `DBSCAN.fit_predict` do when chained, minus
the Python objects in between.  With torch.distributed initialised (one process per GPU) the rows
are partitioned across ranks -- see sharded.py.
"""
import numpy as np
import torch

from . import _lib
from ._lib import call, ptr, stream_ptr
from .dbscan import dbscan_from_neighbors
from .faiss_rerank import jaccard_neighbors, rerank_state


@torch.no_grad()
def pseudo_labels(x, k1=30, k2=6, eps=0.6, min_samples=4, knn="auto", centroids=False, timers=False):
    """x: (N, D) fp32 CUDA tensor, rows L2-normalised.  Returns dict(labels int64 cuda (N,), core uint8 cuda,
    num_clusters 1-elem int64 cuda, [centroids (C, D) cuda], state)."""
    if not x.is_cuda:
        raise RuntimeError("pseudo_labels needs a CUDA tensor; there is no CPU fallback")
    with torch.cuda.device(x.device):
        st = rerank_state(x.contiguous(), k1, k2, knn=knn, timers=timers)
        ev0 = ev1 = ev2 = None
        if timers:
            ev0 = torch.cuda.Event(enable_timing=True)
            ev0.record()
        slot_ptr, nbr_idx, nbr_cnt, _ = jaccard_neighbors(st, eps)
        if timers:
            ev1 = torch.cuda.Event(enable_timing=True)
            ev1.record()
        labels, core, ncl = dbscan_from_neighbors(st.N, slot_ptr, nbr_idx, nbr_cnt, min_samples)
        if timers:
            ev2 = torch.cuda.Event(enable_timing=True)
            ev2.record()
            torch.cuda.synchronize()
            st.timings["jaccard"] = ev0.elapsed_time(ev1) * 1e-3
            st.timings["dbscan"] = ev1.elapsed_time(ev2) * 1e-3
        out = dict(labels=labels, core=core, num_clusters=ncl, state=st, nbr_cnt=nbr_cnt)
        if centroids:
            C = int(ncl.item())
            cen = torch.empty((C, x.shape[1]), dtype=torch.float32, device=x.device)
            if C:
                call("reid_centroids", ptr(x), x.shape[0], x.shape[1], ptr(labels), C, 1, ptr(cen), None, stream_ptr())
            out["centroids"] = cen
        return out
