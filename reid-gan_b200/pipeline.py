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
from .faiss_rerank import jaccard_neighbors, rerank_state_async


def _labels_from_state(st, eps, min_samples, timers=False):
    """a7 + a8 on a state (speculative sizes while its report is pending)."""
    ev0 = ev1 = ev2 = None
    if timers:
        ev0 = torch.cuda.Event(enable_timing=True)
        ev0.record()
    # owned pairs: every edge of the eps-graph is accumulated and listed once (J is bit-symmetric)
    slot_ptr, nbr_idx, nbr_cnt, _ = jaccard_neighbors(st, eps, owned=True)
    if timers:
        ev1 = torch.cuda.Event(enable_timing=True)
        ev1.record()
    labels, core, ncl = dbscan_from_neighbors(st.N, slot_ptr, nbr_idx, nbr_cnt, min_samples, owned=True)
    if timers:
        ev2 = torch.cuda.Event(enable_timing=True)
        ev2.record()
        torch.cuda.synchronize()
        st.timings["jaccard"] = ev0.elapsed_time(ev1) * 1e-3
        st.timings["dbscan"] = ev1.elapsed_time(ev2) * 1e-3
    return labels, core, ncl, nbr_cnt


class PassGraph:
    """One whole pseudo-label pass captured in a CUDA graph: ~40 launches (plus the collectives of a row-sharded pass)
    replayed with one cudaGraphLaunch.  Possible because the speculative pass never reads a size back: every buffer is
    an upper bound or a guarded guess, so the launch sequence does not depend on the data; what did not fit is in the
    pass report, which is read after the replay (the caller then redoes the pass eagerly with exact sizes).
    Capture happens on a private stream after one eager pass on that stream (library-owned scratch, tile lists and
    NCCL channels must exist before a capture starts).  The outputs live in the graph's memory pool and are
    overwritten by the next replay."""

    _cache = {}

    def __init__(self, fn):
        self.stream = torch.cuda.Stream()
        self.graph = torch.cuda.CUDAGraph()
        cur = torch.cuda.current_stream()
        self.stream.wait_stream(cur)
        with torch.cuda.stream(self.stream):
            fn()                                         # warm-up on the capture stream
        self.stream.synchronize()
        with torch.cuda.graph(self.graph, stream=self.stream, capture_error_mode="thread_local"):
            self.result = fn()
        cur.wait_stream(self.stream)

    @classmethod
    def get(cls, key, fn):
        g = cls._cache.get(key)
        if g is None:
            if len(cls._cache) >= 2:                     # a graph pins its buffers (hundreds of MB): keep at most two
                cls._cache.clear()
            g = cls._cache[key] = PassGraph(fn)
        return g


def _graph_key(x, *params):
    return (x.data_ptr(), tuple(x.shape), x.device.index) + tuple(params)


@torch.no_grad()
def pseudo_labels(x, k1=30, k2=6, eps=0.6, min_samples=4, knn="auto", centroids=False, timers=False, graph=False):
    """x: (N, D) fp32 CUDA tensor, rows L2-normalised.  Returns dict(labels int64 cuda (N,), core uint8 cuda,
    num_clusters 1-elem int64 cuda, [centroids (C, D) cuda], state).

    The pass is enqueued without a single host synchronisation (speculative sizes, faiss_rerank.py) and ends with ONE
    read-back: the pass report and, for the centroids, the number of clusters.  graph=True replays the pass from a CUDA
    graph captured at the first call with this tensor / these parameters (PassGraph)."""
    if not x.is_cuda:
        raise RuntimeError("pseudo_labels needs a CUDA tensor; there is no CPU fallback")
    with torch.cuda.device(x.device):
        x = x.contiguous()
        cap = max(1, x.shape[0])

        def enqueue():
            st = rerank_state_async(x, k1, k2, knn=knn, timers=timers)
            labels, core, ncl, nbr_cnt = _labels_from_state(st, eps, min_samples, timers)
            cen = None
            if centroids:
                # C is not known on the host yet (at most N: every cluster holds a core point): the kernel takes the
                # capacity and the device-side count, rows beyond the count are never written (nor their memory touched)
                cen = torch.empty((cap, x.shape[1]), dtype=torch.float32, device=x.device)
                cws = torch.empty(max(1, _lib.lib().reid_centroids_workspace_bytes(x.shape[0], cap)), dtype=torch.uint8, device=x.device)
                call("reid_centroids_dev", ptr(x), x.shape[0], x.shape[1], ptr(labels), ptr(ncl), cap, 1, ptr(cen), ptr(cws), stream_ptr())
            return st, labels, core, ncl, nbr_cnt, cen, st.report

        if graph and not timers:
            from .faiss_rerank import partner_guess_on          # a changed hint changes the launch sequence: new graph
            g = PassGraph.get(_graph_key(x, k1, k2, float(eps), min_samples, knn, bool(centroids), partner_guess_on(x.shape[0])),
                              enqueue)
            g.graph.replay()
            st, labels, core, ncl, nbr_cnt, cen, report = g.result
            st.report = report                              # finish() consumes it: hand the (refilled) tensor back
        else:
            st, labels, core, ncl, nbr_cnt, cen, _ = enqueue()
        st2, nbr_ok = st.finish(check_nbr=True)
        if st2 is not st or not nbr_ok:                       # a guess did not hold: exact sizes (rare)
            st = st2
            labels, core, ncl, nbr_cnt = _labels_from_state(st, eps, min_samples, timers)
            if centroids:
                cen = torch.empty((cap, x.shape[1]), dtype=torch.float32, device=x.device)
                cws = torch.empty(max(1, _lib.lib().reid_centroids_workspace_bytes(x.shape[0], cap)), dtype=torch.uint8, device=x.device)
                call("reid_centroids_dev", ptr(x), x.shape[0], x.shape[1], ptr(labels), ptr(ncl), cap, 1, ptr(cen), ptr(cws), stream_ptr())
        out = dict(labels=labels, core=core, num_clusters=ncl, state=st, nbr_cnt=nbr_cnt)
        if centroids:
            out["centroids"] = cen[: int(ncl.item())]
        return out
