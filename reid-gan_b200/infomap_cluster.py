"""Drop-in for the kNN front end of the Infomap clustering variant, clustercontrast/utils/infomap_cluster.py
(the path examples/cluster_contrast_train_usl_infomap.py:169-173 runs):

    feat_dists, feat_nbrs = get_dist_nbr(features=features_array, k=args.k1, knn_method='faiss-gpu')   # :230-234
    single, links = get_links(single=[], links={}, nbrs=feat_nbrs, dists=feat_dists, min_sim=eps)      # :129-144

`get_dist_nbr` is the same tensor-core search as the DBSCAN variant's (inner product on unit-norm rows: the canonical
key IS the similarity), `get_links` is a count/scan/fill pair of kernels.  Infomap itself (the third-party `infomap`
package driven by cluster_by_infomap :147-227) is out of scope: it consumes the `links` dict these functions return.
"""
import numpy as np
import torch

from . import _lib
from ._lib import call, ptr, stream_ptr
from .faiss_rerank import _device_of, _scan, knn_search


def get_dist_nbr(features, k=80, knn_method='faiss-cpu', return_device=False):
    """(dists float64 (N, k) ascending = 1 - similarity, nbrs int32 (N, k)); `knn_method` is accepted and ignored
    (one implementation).  With return_device=True the CUDA tensors (float32 dists, int32 nbrs) come back instead."""
    if isinstance(features, np.ndarray):
        features = torch.from_numpy(np.ascontiguousarray(features.astype('float32')))      # :59
    dev = _device_of(features)
    with torch.cuda.device(dev), torch.no_grad():
        x = features.to(device=dev, dtype=torch.float32).contiguous()
        if k > x.shape[0]:
            raise ValueError("k=%d exceeds the number of samples N=%d" % (k, x.shape[0]))
        nbrs, sims, _ = knn_search(x, k, "auto", metric="ip")      # faiss.IndexFlatIP (:70-73): inner product whatever the norms
        dists = 1.0 - sims                                   # :75, fp32
        if return_device:
            return dists, nbrs
        # np.array(knns) at :116 upcasts the (int32, float32) tuples to float64: hand out the same dtypes
        return dists.cpu().numpy().astype(np.float64), nbrs.cpu().numpy().astype(np.int32)


def links_device(nbrs, dists, min_sim):
    """CSR of the links on the device: (ptr int64 (N+1), dst int32, weight float64, single bool (N,))."""
    N, k = nbrs.shape
    dev = nbrs.device
    sp = stream_ptr()
    cnt = torch.empty(N, dtype=torch.int32, device=dev)
    call("reid_links_count", ptr(nbrs), ptr(dists), N, k, float(min_sim), ptr(cnt), sp)
    lp, total, _ = _scan(cnt, N, dev)
    dst = torch.empty(max(total, 1), dtype=torch.int32, device=dev)
    w = torch.empty(max(total, 1), dtype=torch.float64, device=dev)
    call("reid_links_fill", ptr(nbrs), ptr(dists), N, k, float(min_sim), ptr(lp), ptr(dst), ptr(w), sp)
    return lp, dst[:total], w[:total], cnt == 0


def get_links(single, links, nbrs, dists, min_sim):
    """infomap_cluster.py:129-144 with the reference's signature: fills `links` {(i, j): similarity} and `single`."""
    dev = _device_of(nbrs if isinstance(nbrs, torch.Tensor) else None)
    with torch.cuda.device(dev), torch.no_grad():
        n_d = torch.as_tensor(np.ascontiguousarray(nbrs) if isinstance(nbrs, np.ndarray) else nbrs).to(dev, torch.int32).contiguous()
        d_d = torch.as_tensor(np.ascontiguousarray(dists) if isinstance(dists, np.ndarray) else dists).to(dev, torch.float32).contiguous()
        lp, dst, w, is_single = links_device(n_d, d_d, min_sim)
        lp_h, dst_h, w_h = lp.cpu().numpy(), dst.cpu().numpy(), w.cpu().numpy()
        src_h = np.repeat(np.arange(n_d.shape[0]), np.diff(lp_h))
        links.update(zip(zip(src_h.tolist(), dst_h.tolist()), w_h.tolist()))
        single.extend(np.nonzero(is_single.cpu().numpy())[0].tolist())
    return single, links
