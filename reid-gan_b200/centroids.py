"""Drop-in for the centroid initialisation of the reference's epoch loop:

    cluster_features = generate_cluster_features(pseudo_labels, features)   # train_usl.py:169-184
    memory.features = F.normalize(cluster_features, dim=1).cuda()           # train_usl.py:191

`generate_cluster_features(labels, features)` keeps the closure's signature and returns the
(C, D) per-label means in ascending label order, label -1 skipped -- as a CUDA tensor, since
the only consumer moves it to the GPU anyway.  `normalize=True` fuses the F.normalize of :191.
"""
import numpy as np
import torch

from . import _lib
from ._lib import call, check, ptr, stream_ptr
from .faiss_rerank import _device_of


@torch.no_grad()
def generate_cluster_features(labels, features, normalize=False, num_clusters=None):
    L = _lib.lib()
    dev = _device_of(features)
    with torch.cuda.device(dev):
        x = features.to(device=dev, dtype=torch.float32).contiguous()
        if isinstance(labels, np.ndarray):
            labels = torch.from_numpy(labels.astype(np.int64, copy=False))
        lab = torch.as_tensor(labels).to(device=dev, dtype=torch.int64).contiguous()
        N, D = x.shape
        if lab.numel() != N:
            raise ValueError("labels (%d) and features (%d) disagree" % (lab.numel(), N))
        if num_clusters is None:
            # the reference keeps only labels that occur (sorted(centers.keys())); DBSCAN labels are
            # dense 0..C-1, so C = max + 1
            num_clusters = int(lab.max().item()) + 1 if N else 0
        C = int(num_clusters)
        if C <= 0:
            raise RuntimeError("generate_cluster_features: no clusters (all labels are -1); "
                               "the reference fails at torch.stack of an empty list here too")
        out = torch.empty((C, D), dtype=torch.float32, device=dev)
        ws = torch.empty(max(1, L.reid_centroids_workspace_bytes(N, C)), dtype=torch.uint8, device=dev)
        call("reid_centroids", ptr(x), N, D, ptr(lab), C, 1 if normalize else 0, ptr(out), ptr(ws), stream_ptr())
        return out
