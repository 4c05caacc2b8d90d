"""a1 on the tensor cores: fp16 tcgen05 candidate GEMM with fused top-K, exact re-score with a
per-row certificate, exact CUDA-core search for the rows that cannot be certified
(replaces faiss IndexFlatL2 / GpuIndexFlatL2.search, utils/faiss_rerank.py:39-62).

The result is bit-identical to the all-exact search whatever the fp16 arithmetic does: a row is
accepted only when the certificate of knn_rescore.cu holds, otherwise it is recomputed exactly.
"""
import ctypes

import torch

from . import _lib
from ._lib import call, check, ptr, stream_ptr

TC_CAP = 512
KEEP_MAX = 64
SCALE_LOG2 = 4          # features are scaled by 2^4 before rounding to fp16 (keeps them out of the subnormals)
CTA_GROUP = int(__import__("os").environ.get("REID_TC_CTA_GROUP", "2"))   # 2 = CTA pairs (tcgen05 cta_group::2)
ORDER_ROWS = __import__("os").environ.get("REID_RESCORE_ORDER", "1") != "0"   # cluster-locality visiting order
SLACK = 34              # K = k + SLACK candidates are kept per (row, column range)


def err_bound(max_sqnorm):
    """|fp16-GEMM score - exact dot| <= 2^-10 ||x_i|| ||x_j|| (both operands rounded to 11 significant
    bits, Cauchy-Schwarz) + fp32 accumulation allowance."""
    return float(max_sqnorm) * (2.0 ** -10) * 1.02 + 2.0 ** -14


def knn_search_tc(x, k, r0, r1, idx, key, info, xh=None, max_sqnorm=None):
    L = _lib.lib()
    from .faiss_rerank import _knn_exact_rows
    N, D = x.shape
    dev = x.device
    n = r1 - r0
    sp = stream_ptr()
    if xh is None:
        xh = torch.empty((N, D), dtype=torch.float16, device=dev)
        msq = torch.zeros(1, dtype=torch.float32, device=dev)
        call("reid_features_to_half", ptr(x), N, D, SCALE_LOG2, ptr(xh), ptr(msq), sp)
        max_sqnorm = None
    else:
        msq = None
    n_splits = ctypes.c_int(1)
    call("reid_knn_tc_plan", N, n, CTA_GROUP, ctypes.byref(n_splits))
    s = n_splits.value
    keep = max(k, min(k + SLACK, KEEP_MAX))
    n_lists = 2 * s                                        # two epilogue groups per column range
    cand = torch.empty(n * n_lists * TC_CAP, dtype=torch.int64, device=dev)
    cand_cnt = torch.zeros(n * n_lists, dtype=torch.int32, device=dev)
    row_tau = torch.empty(n, dtype=torch.int32, device=dev)
    call("reid_knn_candidates_tc", ptr(xh), N, D, SCALE_LOG2, r0, r1, keep, s, CTA_GROUP, ptr(cand), ptr(cand_cnt), ptr(row_tau), sp)
    eps = err_bound(max_sqnorm) if max_sqnorm is not None else 0.0   # else derived on device from msq
    flag = torch.empty(n, dtype=torch.int32, device=dev)
    max_err = torch.zeros(1, dtype=torch.float32, device=dev)
    ws = torch.empty(L.reid_knn_rescore_workspace_bytes(N, n), dtype=torch.uint8, device=dev)
    call("reid_knn_rescore", ptr(x), N, D, r0, r1, ptr(cand), ptr(cand_cnt), ptr(row_tau), n_lists, k, eps,
         ptr(msq), 1 if ORDER_ROWS else 0, ptr(idx), ptr(key), ptr(flag), ptr(max_err), ptr(ws), sp)
    bad = torch.nonzero(flag).flatten().to(torch.int32)
    n_bad = bad.numel()
    if n_bad:
        rows = (bad + r0).contiguous()
        bi = torch.empty((n_bad, k), dtype=torch.int32, device=dev)
        bk = torch.empty((n_bad, k), dtype=torch.float32, device=dev)
        _knn_exact_rows(x, k, rows, 0, n_bad, bi, bk)
        idx[bad.long()] = bi
        key[bad.long()] = bk
    info.update(mode="tc", cta_group=CTA_GROUP, n_splits=s, keep=keep, err_bound=eps if msq is None else None, max_sqnorm=msq, uncertified_rows=int(n_bad),
                max_abs_err=max_err, xh=xh)
    return idx, key, info
