"""Drop-in for clustercontrast/utils/rerank.py:

    distmat = re_ranking(distmat.numpy(), distmat_qq.numpy(), distmat_gg.numpy())     # evaluators.py:141
    re_ranking(q_g_dist, q_q_dist, g_g_dist, k1=20, k2=6, lambda_value=0.3) -> float32 (Q, G)

the evaluation-time k-reciprocal re-ranking behind Evaluator.evaluate(rerank=True).  Same mathematics as the
pseudo-label path on a (query+gallery)^2 matrix of column-max-normalised squared distances (rerank.py:36-41), with
exp(-d) weights (:66-67) and the blend (1 - lambda) * Jaccard + lambda * original (:95-96).  The set / sparse / Jaccard
stages are the kernels of the pseudo-label path; csrc/eval_rerank.cu holds the dense front and back ends.
Everything runs on the current CUDA device through libreid_b200.so; there is no CPU fallback.
"""
import numpy as np
import torch

from . import _lib
from ._lib import call, ptr, stream_ptr
from .faiss_rerank import _device_of, _scan, half_k


def _dev_f32(a, dev):
    if isinstance(a, np.ndarray):
        a = torch.from_numpy(np.ascontiguousarray(a))
    return a.to(device=dev, dtype=torch.float32).contiguous()


def re_ranking(q_g_dist, q_q_dist, g_g_dist, k1=20, k2=6, lambda_value=0.3, return_device=False):
    L = _lib.lib()
    dev = _device_of(q_g_dist if isinstance(q_g_dist, torch.Tensor) else None)
    with torch.cuda.device(dev), torch.no_grad():
        qg, qq, gg = _dev_f32(q_g_dist, dev), _dev_f32(q_q_dist, dev), _dev_f32(g_g_dist, dev)
        Q, G = qg.shape
        if qq.shape != (Q, Q) or gg.shape != (G, G):
            raise ValueError("expected q_g (Q, G), q_q (Q, Q), g_g (G, G); got %s %s %s" % (tuple(qg.shape), tuple(qq.shape), tuple(gg.shape)))
        M = Q + G
        cols = k1 + 1                                            # initial_rank[i, :k1+1]   (rerank.py:51-52)
        if not (1 <= k1 and cols <= 64 and cols <= M):
            raise ValueError("k1=%d outside the supported range (k1 + 1 <= min(64, Q + G))" % k1)
        if not (1 <= k2 <= cols):
            raise ValueError("k2=%d must be in 1..k1+1" % k2)
        sp = stream_ptr()
        # :36-41  normalised squared distances, transposed
        scratch = torch.empty((M, M), dtype=torch.float32, device=dev)
        colmax = torch.empty(M, dtype=torch.float32, device=dev)
        dist = torch.empty((M, M), dtype=torch.float32, device=dev)
        call("reid_rr_normalised_distance", ptr(qg), ptr(qq), ptr(gg), Q, G, ptr(scratch), ptr(colmax), ptr(dist), sp)
        del scratch
        # :43  initial_rank (only the first k1+1 columns are ever read)
        rank = torch.empty((M, cols), dtype=torch.int32, device=dev)
        call("reid_select_rows", ptr(dist), M, M, cols, 1, ptr(rank), None, sp)
        # :50-64  k-reciprocal sets and their expansion
        h = half_k(k1)
        R = torch.empty(M, dtype=torch.int64, device=dev)
        Rh = torch.empty(M, dtype=torch.int64, device=dev)
        call("reid_reciprocal_masks", ptr(rank), M, cols, k1, 0, M, ptr(R), sp)
        call("reid_reciprocal_masks", ptr(rank), M, cols, h, 0, M, ptr(Rh), sp)
        half_cols = min(h + 1, cols)
        e_stride = min(cols + cols * half_cols, 1024)
        e_pad = torch.empty(M * e_stride, dtype=torch.int32, device=dev)
        e_cnt = torch.empty(M, dtype=torch.int32, device=dev)
        call("reid_expand", ptr(rank), M, cols, half_cols, ptr(R), ptr(Rh), 0, M, e_stride, ptr(e_pad), ptr(e_cnt), sp)
        e_ptr, e_total, e_max = _scan(e_cnt, M, dev)
        if e_max > e_stride:
            raise RuntimeError("re_ranking: an expansion set exceeded %d entries" % e_stride)
        # :66-67  weights
        e_idx = torch.empty(max(e_total, 1), dtype=torch.int32, device=dev)
        v_val = torch.empty(max(e_total, 1), dtype=torch.float32, device=dev)
        call("reid_rr_weights", ptr(dist), M, ptr(e_pad), e_stride, ptr(e_ptr), M, ptr(e_idx), ptr(v_val), sp)
        # :69-74  query expansion
        if k2 != 1:
            q_stride = L.reid_query_expand_stride(k2, max(e_max, 1))
            q_cnt = torch.empty(M, dtype=torch.int32, device=dev)
            qp_idx = torch.empty(M * q_stride, dtype=torch.int32, device=dev)
            qp_val = torch.empty(M * q_stride, dtype=torch.float32, device=dev)
            call("reid_query_expand", ptr(rank), M, cols, k2, ptr(e_ptr), ptr(e_idx), ptr(v_val), max(e_max, 1), 0, M,
                 ptr(q_cnt), ptr(qp_idx), ptr(qp_val), None, 0, sp)
            q_ptr, q_total, _ = _scan(q_cnt, M, dev)
            q_idx = torch.empty(max(q_total, 1), dtype=torch.int32, device=dev)
            q_val = torch.empty(max(q_total, 1), dtype=torch.float32, device=dev)
            call("reid_csr_compact", ptr(qp_idx), ptr(qp_val), q_stride, ptr(q_cnt), ptr(q_ptr), M, ptr(q_idx), ptr(q_val), sp)
        else:
            q_ptr, q_idx, q_val, q_total = e_ptr, e_idx, v_val, e_total
        # :76-78  inverted index
        c_cnt = torch.empty(M, dtype=torch.int32, device=dev)
        call("reid_transpose_count", ptr(q_idx), q_total, None, M, ptr(c_cnt), sp)
        c_ptr, _, c_max = _scan(c_cnt, M, dev)
        c_idx = torch.empty(max(q_total, 1), dtype=torch.int32, device=dev)
        c_val = torch.empty(max(q_total, 1), dtype=torch.float32, device=dev)
        call("reid_transpose_fill", ptr(q_ptr), ptr(q_idx), ptr(q_val), M, M, ptr(c_ptr), ptr(c_cnt), ptr(c_idx), ptr(c_val),
             int(c_max), sp)
        # :80-93  Jaccard rows of the queries, in blocks of rows; :95-96 blend and slice
        out = torch.empty((Q, G), dtype=torch.float32, device=dev)
        block = max(1, min(Q, (512 << 20) // (4 * M)))
        a_ = float(np.float32(1 - lambda_value))
        b_ = float(np.float32(lambda_value))
        for r0 in range(0, Q, block):
            r1 = min(Q, r0 + block)
            J = torch.empty((r1 - r0, M), dtype=torch.float32, device=dev)
            call("reid_jaccard_dense", ptr(q_ptr), ptr(q_idx), ptr(q_val), ptr(c_ptr), ptr(c_idx), ptr(c_val), M, r0, r1, ptr(J),
                 M, 0, sp)
            call("reid_rr_final", ptr(J), M, ptr(dist[r0:]), r1 - r0, G, a_, b_, ptr(out[r0:]), sp)
        if return_device:
            return out
        return out.cpu().numpy()
