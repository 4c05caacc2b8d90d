"""Drop-ins for the evaluation metrics:

    pairwise_distance(features, query, gallery)            clustercontrast/evaluators.py:71-88
    mean_ap(distmat, query_ids, gallery_ids, query_cams, gallery_cams)      evaluation_metrics/ranking.py:82-115
    cmc(distmat, ..., topk=100, separate_camera_set, single_gallery_shot, first_match_break)        ranking.py:18-79

The reference argsorts every query row and then only looks at where the positives landed; csrc/eval_metrics.cu
counts, per positive, the valid gallery items at or before it instead -- no sort, one CTA per query.  Ties in
distance are ranked by gallery index (np.argsort at ranking.py:41,104 leaves them unordered); average precision
treats tied scores as one threshold, exactly like sklearn's average_precision_score.
"""
from collections import OrderedDict  # noqa: F401  (the reference passes an OrderedDict of features)

import numpy as np
import torch

from . import _lib
from ._lib import call, ptr, stream_ptr
from .faiss_rerank import _device_of


def pairwise_distance_device(x, y):
    """(m, D), (n, D) float32 CUDA tensors -> (m, n) float32 CUDA tensor  ||x||^2 + ||y||^2 - 2 x.y^T."""
    m, D = x.shape
    n = y.shape[0]
    out = torch.empty((m, n), dtype=torch.float32, device=x.device)
    norms = torch.empty(m + n, dtype=torch.float32, device=x.device)
    call("reid_pairwise_distance", ptr(x), ptr(y), m, n, D, ptr(norms), ptr(out), stream_ptr())
    return out


def pairwise_distance(features, query=None, gallery=None):
    """evaluators.py:71-88.  `features`: dict name -> (D,) tensor.  Without query/gallery: the (n, n) matrix of all
    features (:72-78); else (dist_m (m, n) CPU tensor, x.numpy(), y.numpy()) like the reference."""
    dev = _device_of(None)
    with torch.cuda.device(dev), torch.no_grad():
        if query is None and gallery is None:
            n = len(features)
            x = torch.cat(list(features.values())).view(n, -1)
            xd = x.to(dev, torch.float32).contiguous()
            return pairwise_distance_device(xd, xd).cpu()
        x = torch.cat([features[f].unsqueeze(0) for f, _, _ in query], 0)
        y = torch.cat([features[f].unsqueeze(0) for f, _, _ in gallery], 0)
        m, n = x.size(0), y.size(0)
        x = x.view(m, -1)
        y = y.view(n, -1)
        d = pairwise_distance_device(x.to(dev, torch.float32).contiguous(), y.to(dev, torch.float32).contiguous())
        return d.cpu(), x.cpu().numpy(), y.cpu().numpy()


def _ids(a, n, default, dev):
    if a is None:
        a = default(n)
    return torch.as_tensor(np.asarray(a)).to(dev, torch.int64).contiguous()


def _rank_metrics(distmat, query_ids, gallery_ids, query_cams, gallery_cams, separate_camera_set, topk, first_match_break):
    L = _lib.lib()
    dev = _device_of(distmat if isinstance(distmat, torch.Tensor) else None)
    with torch.cuda.device(dev), torch.no_grad():
        d = torch.as_tensor(distmat if isinstance(distmat, torch.Tensor) else np.ascontiguousarray(distmat))
        d = d.to(dev, torch.float32).contiguous()
        m, n = d.shape
        qi = _ids(query_ids, m, np.arange, dev)                                        # ranking.py:26-33 / 88-95
        gi = _ids(gallery_ids, n, np.arange, dev)
        qc = _ids(query_cams, m, lambda k: np.zeros(k).astype(np.int32), dev)
        gc = _ids(gallery_cams, n, lambda k: np.ones(k).astype(np.int32), dev)
        ap = torch.empty(m, dtype=torch.float64, device=dev)
        has = torch.empty(m, dtype=torch.int32, device=dev)
        contrib = ret = None
        if topk:
            contrib = torch.empty((m, topk), dtype=torch.float64, device=dev)
            ret = torch.empty(topk, dtype=torch.float64, device=dev)
        call("reid_rank_metrics", ptr(d), m, n, n, ptr(qi), ptr(gi), ptr(qc), ptr(gc), int(bool(separate_camera_set)),
             int(topk), int(bool(first_match_break)), ptr(ap), ptr(has), ptr(contrib), ptr(ret), stream_ptr())
        has_h = has.cpu().numpy()
        if (has_h < 0).any():
            raise RuntimeError("a query has more positives than reid_rank_metrics handles")
        return ap.cpu().numpy(), has_h, None if ret is None else ret.cpu().numpy()


def mean_ap(distmat, query_ids=None, gallery_ids=None, query_cams=None, gallery_cams=None):
    ap, has, _ = _rank_metrics(distmat, query_ids, gallery_ids, query_cams, gallery_cams, False, 0, False)
    aps = ap[has > 0]
    if len(aps) == 0:
        raise RuntimeError("No valid query")
    return np.mean(aps)


def cmc(distmat, query_ids=None, gallery_ids=None, query_cams=None, gallery_cams=None, topk=100,
        separate_camera_set=False, single_gallery_shot=False, first_match_break=False):
    if single_gallery_shot:
        return _cmc_single_gallery_shot(distmat, query_ids, gallery_ids, query_cams, gallery_cams, topk,
                                        separate_camera_set, first_match_break)
    _, has, ret = _rank_metrics(distmat, query_ids, gallery_ids, query_cams, gallery_cams, separate_camera_set, topk,
                                first_match_break)
    num_valid_queries = int((has > 0).sum())
    if num_valid_queries == 0:
        raise RuntimeError("No valid query")
    return ret.cumsum() / num_valid_queries


def _cmc_single_gallery_shot(distmat, query_ids, gallery_ids, query_cams, gallery_cams, topk, separate_camera_set,
                             first_match_break):
    """ranking.py:53-66: per valid query, 10 times, ONE gallery instance per identity is drawn with np.random.choice
    (ranking.py:10-16) and the CMC contribution is computed on that subset.  The draws consume the process-wide
    np.random state in a fixed protocol -- query ascending, repeat, identities in order of first appearance in the
    ranking -- so this option is a host-side twin by definition: same seed, same draws, same result as the
    reference.  Only the ranking itself comes from the device (rows sorted by distance, ties by gallery index; the
    reference's np.argsort leaves ties unordered)."""
    from collections import defaultdict
    dev = _device_of(distmat if isinstance(distmat, torch.Tensor) else None)
    with torch.cuda.device(dev), torch.no_grad():
        d = torch.as_tensor(distmat if isinstance(distmat, torch.Tensor) else np.ascontiguousarray(distmat))
        d = d.to(dev, torch.float32)
        m, n = d.shape
        indices = torch.sort(d, dim=1, stable=True).indices.cpu().numpy()
    query_ids = np.arange(m) if query_ids is None else np.asarray(query_ids)
    gallery_ids = np.arange(n) if gallery_ids is None else np.asarray(gallery_ids)
    query_cams = np.zeros(m).astype(np.int32) if query_cams is None else np.asarray(query_cams)
    gallery_cams = np.ones(n).astype(np.int32) if gallery_cams is None else np.asarray(gallery_cams)
    matches = (gallery_ids[indices] == query_ids[:, np.newaxis])
    ret = np.zeros(topk)
    num_valid_queries = 0
    repeat = 10
    for i in range(m):
        valid = ((gallery_ids[indices[i]] != query_ids[i]) | (gallery_cams[indices[i]] != query_cams[i]))
        if separate_camera_set:
            valid &= (gallery_cams[indices[i]] != query_cams[i])
        if not np.any(matches[i, valid]):
            continue
        gids = gallery_ids[indices[i][valid]]
        inds = np.where(valid)[0]
        ids_dict = defaultdict(list)
        for j, x in zip(inds, gids):
            ids_dict[x].append(j)
        for _ in range(repeat):
            mask = np.zeros(len(valid), dtype=bool)
            for _, idx_list in ids_dict.items():
                mask[np.random.choice(idx_list)] = True
            sampled = valid & mask
            index = np.nonzero(matches[i, sampled])[0]
            delta = 1. / (len(index) * repeat)
            for j, k in enumerate(index):
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
