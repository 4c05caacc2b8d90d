"""Drop-in for the DBSCAN step of the reference's pseudo-label pass:

    cluster = DBSCAN(eps=eps, min_samples=4, metric='precomputed', n_jobs=-1)
    pseudo_labels = cluster.fit_predict(rerank_dist)
        -- examples/cluster_contrast_train_usl.py:160,163 (sklearn.cluster.DBSCAN)

Only metric='precomputed' exists here (the only use on the path).  `fit_predict` accepts
  * the `JaccardDistance` returned by reid_gan_b200.compute_jaccard_distance: the eps-graph is
    built straight from the device-resident sparse V_qe (no N x N matrix anywhere), or
  * a dense (N, N) float matrix (numpy or torch, host or device) like sklearn does.
Labels follow sklearn exactly: cluster ids by ascending smallest core index, border points
take the smallest adjacent core label, noise is -1; returned as np.intp.
"""
import numpy as np
import torch

from . import _lib
from ._lib import call, check, ptr, stream_ptr
from .faiss_rerank import JaccardDistance, REPORT_WORDS, R_ESC, R_NBR_OVF, R_S, _guess_hint, _device_of, _nbr_cap_hint, _scan, jaccard_neighbors


def dbscan_from_neighbors(N, nbr_ptr, nbr_idx, nbr_cnt, min_samples, owned=False):
    """Device labelling from eps-neighbour lists of all N rows -> (labels int64 cuda, core uint8 cuda, n_clusters).
    owned: the lists name every edge once (jaccard_neighbors(owned=True))."""
    L = _lib.lib()
    dev = nbr_ptr.device
    labels = torch.empty(N, dtype=torch.int64, device=dev)
    core = torch.empty(N, dtype=torch.uint8, device=dev)
    ncl = torch.zeros(1, dtype=torch.int64, device=dev)
    ws = torch.empty(max(1, L.reid_dbscan_workspace_bytes(N)), dtype=torch.uint8, device=dev)
    call("reid_dbscan_labels", N, ptr(nbr_ptr), ptr(nbr_idx), ptr(nbr_cnt), int(min_samples), ptr(labels), ptr(core),
                               ptr(ncl), ptr(ws), 1 if owned else 0, stream_ptr())
    return labels, core, ncl


class DBSCAN:
    def __init__(self, eps=0.5, *, min_samples=5, metric="precomputed", metric_params=None, algorithm="auto",
                 leaf_size=30, p=None, n_jobs=None):
        if metric != "precomputed":
            raise ValueError("reid_gan_b200.DBSCAN only implements metric='precomputed' "
                             "(the reference's call, cluster_contrast_train_usl.py:160)")
        if not eps > 0.0:
            raise ValueError("eps must be positive")
        if int(min_samples) < 1:
            raise ValueError("min_samples must be >= 1")
        self.eps = eps
        self.min_samples = int(min_samples)
        self.metric = metric
        self.metric_params = metric_params
        self.algorithm = algorithm
        self.leaf_size = leaf_size
        self.p = p
        self.n_jobs = n_jobs          # inert for precomputed input in sklearn too (_dbscan.py:122-123)
        self.labels_ = None
        self.core_sample_indices_ = None

    # sklearn API -------------------------------------------------------------------
    def fit(self, X, y=None, sample_weight=None):
        if sample_weight is not None:
            raise NotImplementedError("sample_weight is not supported")
        if isinstance(X, JaccardDistance):
            labels, core = self._fit_sparse(X)
        else:
            labels, core = self._fit_dense(X)
        self.labels_ = labels.cpu().numpy().astype(np.intp, copy=False)
        self.core_sample_indices_ = np.nonzero(core.cpu().numpy())[0].astype(np.intp)
        return self

    def fit_predict(self, X, y=None, sample_weight=None):
        return self.fit(X, sample_weight=sample_weight).labels_

    # device paths ------------------------------------------------------------------
    def _fit_sparse(self, dist):
        st = dist.state
        if st.row_begin != 0 or st.row_end != st.N:
            raise ValueError("DBSCAN needs the full matrix; use sharded.pseudo_labels for row shards")
        with torch.cuda.device(st.Q_ptr.device), torch.no_grad():
            if float(np.float32(self.eps)) >= 1.0:
                # pairs without a shared column have J == 1 and are not in the sparse form
                return self._fit_dense(dist.dense_device())
            # speculative slot sizes first (no host round trip before the labels); the overflow count comes back
            # with the same synchronisation that fetches the labels, and a pass that did not fit is redone exactly
            report = torch.zeros(REPORT_WORDS, dtype=torch.int64, device=st.Q_ptr.device)
            st.report = report
            try:
                slot_ptr, nbr_idx, nbr_cnt, _ = jaccard_neighbors(st, self.eps, speculative=True, owned=True)
            finally:
                st.report = None
            labels, core, _ = dbscan_from_neighbors(st.N, slot_ptr, nbr_idx, nbr_cnt, self.min_samples, owned=True)
            vals = report.tolist()
            if vals[R_NBR_OVF]:
                slot_ptr, nbr_idx, nbr_cnt, _ = jaccard_neighbors(st, self.eps, speculative=False, owned=True)
                labels, core, _ = dbscan_from_neighbors(st.N, slot_ptr, nbr_idx, nbr_cnt, self.min_samples, owned=True)
            elif vals[R_S]:
                _nbr_cap_hint[st.N] = int(vals[R_S])
                if st.N not in _guess_hint and vals[R_ESC] * 4 > 3 * st.N:
                    _guess_hint[st.N] = False
        return labels, core

    def _fit_dense(self, X):
        L = _lib.lib()
        if isinstance(X, np.ndarray):
            X = torch.from_numpy(np.ascontiguousarray(X))
        if not isinstance(X, torch.Tensor):
            X = torch.as_tensor(np.asarray(X))
        if X.dim() != 2 or X.shape[0] != X.shape[1]:
            raise ValueError("precomputed distance matrix must be square, got %s" % (tuple(X.shape),))
        N = X.shape[0]
        dev = _device_of(X)
        # sklearn compares in the matrix dtype: fp32 (or narrower) input -> d <= float32(eps) on the fp32 values; a
        # float64 matrix is compared in float64 (the block is reduced to {0, 2} with that comparison before the fp32
        # kernels see it, so no distance is rounded across eps)
        wide = X.dtype == torch.float64
        eps32 = 1.0 if wide else float(np.float32(self.eps))
        with torch.cuda.device(dev), torch.no_grad():
            sp = stream_ptr()
            block = N if X.is_cuda else max(1, min(N, (512 << 20) // (4 * N)))
            cnts, idxs = [], []
            for a in range(0, N, block):
                b = min(N, a + block)
                if wide:
                    blk = torch.where(X[a:b].to(device=dev, non_blocking=True) <= float(self.eps), 0.0, 2.0).to(torch.float32).contiguous()
                else:
                    blk = X[a:b].to(device=dev, dtype=torch.float32, non_blocking=True).contiguous()
                cnt = torch.empty(b - a, dtype=torch.int32, device=dev)
                call("reid_dbscan_dense_count", ptr(blk), N, blk.stride(0), eps32, 0, b - a, ptr(cnt), sp)
                p_loc, total, _ = _scan(cnt, b - a, dev)
                idx = torch.empty(max(total, 1), dtype=torch.int32, device=dev)
                call("reid_dbscan_dense_fill", ptr(blk), N, blk.stride(0), eps32, 0, b - a, ptr(p_loc), ptr(idx), sp)
                cnts.append(cnt)
                idxs.append(idx[:total])
            cnt = torch.cat(cnts) if len(cnts) > 1 else cnts[0]
            idx = torch.cat(idxs) if len(idxs) > 1 else idxs[0]
            if idx.numel() == 0:
                idx = torch.zeros(1, dtype=torch.int32, device=dev)
            nbr_ptr, _, _ = _scan(cnt, N, dev)
            labels, core, _ = dbscan_from_neighbors(N, nbr_ptr, idx, cnt, self.min_samples)
        return labels, core
