"""ORACLE (test infrastructure, never shipped, never on the product path).

CPU restatement, in numpy, of the k-reciprocal Jaccard re-ranking of
cluster-contrast-reid:  clustercontrast/utils/faiss_rerank.py:23-123
(`k_reciprocal_neigh`, `compute_jaccard_distance`).  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import this module.

Parity status: the reference holds NO tests, golden vectors or fixtures for this
path.  This restatement is instead pinned against the
reference's own Python, executed verbatim in the build container through
`oracle/ref_shim.py`; the outputs are committed under `tests/golden/` together
with the generating script `oracle/make_golden.py`.

The kNN search itself lives in faiss (pip `faiss_gpu`, un-vendored and
un-pinned, reference setup.py:13), absent from /root/reference and from this
image.  Its published contract for IndexFlatL2.search (call site
faiss_rerank.py:58-62) is "exact k smallest squared-L2, ascending"; fp32
rounding inside its SGEMM is not reproducible, so parity is anchored on the
canonical key  fp32( sum_k x_ik*x_jk accumulated in fp64 )  ordered by
(key descending, index ascending)  -- for unit-norm rows identical to L2
ascending.

Two flavours of every stage are kept:
  *_loops   : line-by-line restatement of the reference loops (dense, small N)
  (no suffix): sparse / vectorised restatement used at N = 32,621 and up,
               proven equal to the loop flavour in tests/test_oracle.py.
"""
import numpy as np

__all__ = [
    "half_k", "exact_knn", "rows_have_one_norm", "k_reciprocal_masks", "reciprocal_lists",
    "expand", "expand_loops", "v_weights", "query_expand", "transpose_csr",
    "jaccard_sparse", "jaccard_dense_from_sparse", "jaccard_dense_loops",
    "compute_jaccard_distance_oracle", "compute_jaccard_distance_half", "sparse_pipeline",
]


def half_k(k1):
    """faiss_rerank.py:69 -- int(np.around(k1/2)), round-half-to-even."""
    return int(np.around(k1 / 2))


# --------------------------------------------------------------------------
# a1  kNN search  (faiss_rerank.py:58-62; faiss IndexFlatL2 contract)
# --------------------------------------------------------------------------
UNIT_NORM_TOL = 1e-4      # max ||x||^2 / min ||x||^2 - 1 below which rows count as "all the same norm"


def rows_have_one_norm(x):
    """faiss IndexFlatL2 ranks by squared L2; that is the inner-product order iff every row has the same norm.  Rows
    that a backbone L2-normalised in fp32 differ by ~1e-7 in squared norm -- below what the reference's own fp32 search
    resolves -- so they are searched with the inner-product key; anything beyond UNIT_NORM_TOL gets the L2 key."""
    n = np.einsum("ij,ij->i", x, x, dtype=np.float64)
    return float(n.max()) <= float(n.min()) * (1.0 + UNIT_NORM_TOL)


def exact_knn(x, k, chunk=2048, return_keys=False, rows=None, metric="auto"):
    """Exact top-k by the canonical key.  x: (N, D) float32.  Returns int64 (n, k)
    (and the fp32 keys).  `rows` restricts the query rows (all N columns searched).
    metric "ip": key = fp32(x_i.x_j in fp64);  "l2": key = fp32(x_i.x_j - ||x_j||^2 / 2 in fp64), i.e. squared L2
    ascending with the per-query constant ||x_i||^2 dropped;  "auto": "ip" when rows_have_one_norm(x) else "l2"."""
    x = np.ascontiguousarray(np.asarray(x, dtype=np.float32))
    N = x.shape[0]
    x64 = x.astype(np.float64)
    if metric == "auto":
        metric = "ip" if rows_have_one_norm(x) else "l2"
    half_n = 0.5 * np.einsum("ij,ij->i", x64, x64) if metric == "l2" else None
    q = np.arange(N) if rows is None else np.asarray(rows)
    out = np.empty((q.size, k), dtype=np.int64)
    keys = np.empty((q.size, k), dtype=np.float32)
    for s in range(0, q.size, chunk):
        qq = q[s:s + chunk]
        key = x64[qq] @ x64.T                               # fp64 accumulate, round once
        if half_n is not None:
            key -= half_n[None, :]
        key = key.astype(np.float32)
        # top-k by (key desc, idx asc).  argpartition is arbitrary inside ties,
        # so rows with a tie across the k-th boundary take the slow exact path.
        part = np.argpartition(-key, k - 1, axis=1)[:, :k]
        kth = np.take_along_axis(key, part, 1).min(axis=1)
        n_ge = (key >= kth[:, None]).sum(axis=1)
        for r in range(qq.size):
            if n_ge[r] == k:
                cand = part[r]
            else:
                cand = np.nonzero(key[r] >= kth[r])[0]
            o = np.lexsort((cand, -key[r, cand].astype(np.float64)))[:k]
            out[s + r] = cand[o]
            keys[s + r] = key[r, cand[o]]
    return (out, keys) if return_keys else out


# --------------------------------------------------------------------------
# a2  reciprocal sets  (faiss_rerank.py:23-27, 65-69)
# --------------------------------------------------------------------------
def k_reciprocal_masks(rank, k, chunk=4096):
    """mask[i, r] = (i in rank[rank[i, r], :k+1]) for r < min(k+1, ncols).
    R_k(i) = rank[i, :cols][mask[i]]  (ordered by rank, faiss_rerank.py:24-27)."""
    N, ncols = rank.shape
    cols = min(k + 1, ncols)
    mask = np.empty((N, cols), dtype=bool)
    for s in range(0, N, chunk):
        fwd = rank[s:s + chunk, :cols]
        bwd = rank[fwd][:, :, :cols]
        me = np.arange(s, min(s + chunk, N))[:, None, None]
        mask[s:s + chunk] = (bwd == me).any(axis=2)
    return mask


def reciprocal_lists(rank, k):
    m = k_reciprocal_masks(rank, k)
    cols = m.shape[1]
    return [rank[i, :cols][m[i]] for i in range(rank.shape[0])]


# --------------------------------------------------------------------------
# a3  expansion  (faiss_rerank.py:72-80)
# --------------------------------------------------------------------------
def expand_loops(R, Rh):
    """Line-by-line: E(i) = unique(R(i) ++ every R_half(c), c in R(i), that passes the 2/3 test)."""
    E = []
    for i in range(len(R)):
        k_reciprocal_index = R[i]
        exp = k_reciprocal_index
        for c in k_reciprocal_index:
            cand = Rh[c]
            if len(np.intersect1d(cand, k_reciprocal_index)) > 2 / 3 * len(cand):
                exp = np.append(exp, cand)
        E.append(np.unique(exp))
    return E


def _lists_to_csr(lists, N):
    import scipy.sparse as sp
    cnt = np.fromiter((len(l) for l in lists), dtype=np.int64, count=len(lists))
    indptr = np.concatenate(([0], np.cumsum(cnt)))
    indices = np.concatenate(lists) if len(lists) else np.zeros(0, np.int64)
    data = np.ones(indices.size, dtype=np.int32)
    return sp.csr_matrix((data, indices, indptr), shape=(len(lists), N))


def expand(rank, k1):
    """Vectorised expansion.  Returns (indptr int64 (N+1,), indices int64 sorted per row)."""
    import scipy.sparse as sp
    N = rank.shape[0]
    h = half_k(k1)
    m1 = k_reciprocal_masks(rank, k1)
    mh = k_reciprocal_masks(rank, h)
    ii, rr = np.nonzero(m1)
    Rm = sp.csr_matrix((np.ones(ii.size, np.int32), (ii, rank[ii, rr])), shape=(N, N))
    ih, rh = np.nonzero(mh)
    Rh = sp.csr_matrix((np.ones(ih.size, np.int32), (ih, rank[ih, rh])), shape=(N, N))
    m_half = np.asarray(Rh.sum(axis=1)).ravel()                # |R_half(c)|
    # cnt[i, c] = |R_half(c) & R(i)| on the pattern c in R(i)
    inter = (Rm @ Rh.T).tocsr()
    Rm.sort_indices()
    coo = Rm.tocoo()
    cnt = np.asarray(inter[coo.row, coo.col]).ravel()
    ok = 3 * cnt > 2 * m_half[coo.col]                        # == cnt > 2/3*m for m < 400
    S = sp.csr_matrix((np.ones(int(ok.sum()), np.int32), (coo.row[ok], coo.col[ok])), shape=(N, N))
    Em = (Rm + S @ Rh).tocsr()
    Em.sort_indices()
    return Em.indptr.astype(np.int64), Em.indices.astype(np.int64)


# --------------------------------------------------------------------------
# a4  Gaussian weights  (faiss_rerank.py:81-85)
# --------------------------------------------------------------------------
def v_weights(x, indptr, indices, chunk=16384):
    """V[i, e] = softmax_e(-(2 - 2 x_i.x_e)), fp32 throughout."""
    x = np.asarray(x, dtype=np.float32)
    N = indptr.size - 1
    rows = np.repeat(np.arange(N), np.diff(indptr))
    s = np.empty(indices.size, dtype=np.float32)
    for a in range(0, indices.size, chunk):
        b = min(a + chunk, indices.size)
        s[a:b] = np.einsum("pd,pd->p", x[rows[a:b]], x[indices[a:b]], dtype=np.float32)
    neg = -(np.float32(2) - np.float32(2) * s)                 # -dist, fp32
    vals = np.empty_like(neg)
    nz = np.nonzero(np.diff(indptr))[0]
    starts = indptr[:-1][nz]
    mx = np.maximum.reduceat(neg, starts) if starts.size else np.zeros(0, np.float32)
    e = np.exp(neg - np.repeat(mx, np.diff(indptr)[nz])).astype(np.float32)
    den = np.add.reduceat(e, starts) if starts.size else np.zeros(0, np.float32)
    vals[:] = e / np.repeat(den, np.diff(indptr)[nz])
    return vals.astype(np.float32)


# --------------------------------------------------------------------------
# a5  k2 query expansion  (faiss_rerank.py:89-94)
# --------------------------------------------------------------------------
def query_expand(indptr, indices, vals, rank, k2):
    """V_qe[i] = (V[rank[i,0]] + V[rank[i,1]] + ...) / k2, adds in r order, fp32."""
    import scipy.sparse as sp
    N = indptr.size - 1
    V = sp.csr_matrix((vals.astype(np.float32), indices, indptr), shape=(N, N))
    if k2 == 1:
        return V.indptr.astype(np.int64), V.indices.astype(np.int64), V.data.astype(np.float32)
    acc = V[rank[:, 0]]
    for r in range(1, k2):
        acc = acc + V[rank[:, r]]                              # fp32 a+b on shared entries
        acc = acc.astype(np.float32)
    acc = acc.tocsr()
    acc.sort_indices()
    acc.data = (acc.data / np.float32(k2)).astype(np.float32)
    return acc.indptr.astype(np.int64), acc.indices.astype(np.int64), acc.data


# --------------------------------------------------------------------------
# a6  inverted index  (faiss_rerank.py:98-100)
# --------------------------------------------------------------------------
def transpose_csr(indptr, indices, vals, N):
    import scipy.sparse as sp
    M = sp.csr_matrix((vals, indices, indptr), shape=(indptr.size - 1, N)).tocsc()
    M.sort_indices()
    return M.indptr.astype(np.int64), M.indices.astype(np.int64), M.data.astype(np.float32)


# --------------------------------------------------------------------------
# a7  Jaccard min-sum  (faiss_rerank.py:102-119)
# --------------------------------------------------------------------------
def jaccard_sparse(indptr, indices, vals, N, row_begin=0, row_end=None, max_triples=40_000_000):
    """Sparse t/J for rows [row_begin,row_end): returns CSR (jp, jj, jv) holding
    every pair that shares a column (J < 1 or J == clip 0), J as fp32.
    t_ij accumulates min() over shared columns in ASCENDING column order with
    sequential fp32 adds -- the order of faiss_rerank.py:109-110."""
    n_rows = indptr.size - 1
    row_end = n_rows if row_end is None else row_end
    cp, cr, cv = transpose_csr(indptr, indices, vals, N)
    colcnt = np.diff(cp)
    out_ptr = [0]
    out_j, out_v = [], []
    i = row_begin
    while i < row_end:
        # batch of rows whose triple count stays under max_triples
        tcount = 0
        j = i
        while j < row_end:
            t = int(colcnt[indices[indptr[j]:indptr[j + 1]]].sum())
            if j > i and tcount + t > max_triples:
                break
            tcount += t
            j += 1
        a, b = indptr[i], indptr[j]
        nz_cols = indices[a:b]
        nz_rows = np.repeat(np.arange(i, j), np.diff(indptr[i:j + 1]))
        nz_vals = vals[a:b]
        reps = colcnt[nz_cols]
        tri_i = np.repeat(nz_rows, reps)
        tri_c = np.repeat(nz_cols, reps)
        tri_vi = np.repeat(nz_vals, reps)
        # position of every triple inside its column list
        start = np.repeat(cp[nz_cols], reps)
        within = np.arange(tri_i.size) - np.repeat(np.cumsum(reps) - reps, reps)
        pos = start + within
        tri_j = cr[pos]
        tri_m = np.minimum(tri_vi, cv[pos]).astype(np.float32)
        order = np.lexsort((tri_c, tri_j, tri_i))              # by (i, j, c ascending)
        tri_i, tri_j, tri_m = tri_i[order], tri_j[order], tri_m[order]
        head = np.ones(tri_i.size, dtype=bool)
        head[1:] = (tri_i[1:] != tri_i[:-1]) | (tri_j[1:] != tri_j[:-1])
        hpos = np.nonzero(head)[0]
        runlen = np.diff(np.append(hpos, tri_i.size))
        acc = np.zeros(hpos.size, dtype=np.float32)
        alive = np.arange(hpos.size)
        p = 0
        while alive.size:                                      # strictly sequential fp32 adds
            acc[alive] = acc[alive] + tri_m[hpos[alive] + p]
            p += 1
            alive = alive[runlen[alive] > p]
        J = (np.float32(1) - acc / (np.float32(2) - acc)).astype(np.float32)
        J[J < 0] = np.float32(0)
        pi = tri_i[hpos]
        out_j.append(tri_j[hpos])
        out_v.append(J)
        cnt_rows = np.bincount(pi - i, minlength=j - i)
        for c in cnt_rows:
            out_ptr.append(out_ptr[-1] + int(c))
        i = j
    jj = np.concatenate(out_j) if out_j else np.zeros(0, np.int64)
    jv = np.concatenate(out_v) if out_v else np.zeros(0, np.float32)
    return np.asarray(out_ptr, dtype=np.int64), jj.astype(np.int64), jv.astype(np.float32)


def jaccard_dense_from_sparse(jp, jj, jv, N, row_begin=0):
    n = jp.size - 1
    out = np.ones((n, N), dtype=np.float32)                    # no shared column -> exactly 1.0
    rows = np.repeat(np.arange(n), np.diff(jp))
    out[rows, jj] = jv
    return out


def jaccard_dense_loops(V):
    """Line-by-line restatement of faiss_rerank.py:98-119 on a dense V (small N)."""
    N = V.shape[0]
    invIndex = [np.where(V[:, i] != 0)[0] for i in range(N)]
    jd = np.zeros((N, N), dtype=V.dtype)
    for i in range(N):
        temp_min = np.zeros((1, N), dtype=V.dtype)
        indNonZero = np.where(V[i, :] != 0)[0]
        indImages = [invIndex[ind] for ind in indNonZero]
        for j in range(len(indNonZero)):
            temp_min[0, indImages[j]] = temp_min[0, indImages[j]] + np.minimum(
                V[i, indNonZero[j]], V[indImages[j], indNonZero[j]])
        jd[i] = 1 - temp_min / (2 - temp_min)
    jd[jd < 0] = 0.0
    return jd


# --------------------------------------------------------------------------
# whole path
# --------------------------------------------------------------------------
def sparse_pipeline(x, k1, k2, rank=None):
    """Every intermediate of the path as sparse structures (dict of arrays)."""
    x = np.ascontiguousarray(np.asarray(x, dtype=np.float32))
    N = x.shape[0]
    if rank is None:
        rank = exact_knn(x, k1)
    h = half_k(k1)
    m1 = k_reciprocal_masks(rank, k1)
    mh = k_reciprocal_masks(rank, h)
    ep, ei = expand(rank, k1)
    ev = v_weights(x, ep, ei)
    qp, qi, qv = query_expand(ep, ei, ev, rank, k2)
    return dict(rank=rank, R_mask=m1, Rh_mask=mh, E_ptr=ep, E_idx=ei, V_val=ev,
                Vq_ptr=qp, Vq_idx=qi, Vq_val=qv, N=N)


def compute_jaccard_distance_half(x, k1=20, k2=6, rank=None):
    """use_float16=True (faiss_rerank.py:37): dense restatement for small N.  V, V_qe, temp_min and jaccard_dist are
    float16 arrays; every numpy operation below is the reference's own expression, so numpy applies the same
    float16 roundings (each op evaluated in float32, result rounded to float16)."""
    x = np.ascontiguousarray(np.asarray(x, dtype=np.float32))
    N = x.shape[0]
    if rank is None:
        rank = exact_knn(x, k1)
    ep, ei = expand(rank, k1)
    ev = v_weights(x, ep, ei)                                          # fp32 softmax (:81, F.softmax of fp32)
    V = np.zeros((N, N), dtype=np.float16)                             # :71
    rows = np.repeat(np.arange(N), np.diff(ep))
    V[rows, ei] = ev.astype(np.float16)                                # :83
    if k2 != 1:                                                        # :89-94
        V_qe = np.zeros_like(V, dtype=np.float16)
        for i in range(N):
            V_qe[i, :] = np.mean(V[rank[i, :k2], :], axis=0)
        V = V_qe
    return jaccard_dense_loops(V)                                      # :98-119 line by line, dtype follows V


def compute_jaccard_distance_oracle(x, k1=20, k2=6, rank=None, use_float16=False):
    """Dense (N, N) Jaccard distance -- same contract as faiss_rerank.py:30,123 (float32, or float16 with use_float16)."""
    if use_float16:
        return compute_jaccard_distance_half(x, k1, k2, rank=rank)
    st = sparse_pipeline(x, k1, k2, rank=rank)
    N = st["N"]
    jp, jj, jv = jaccard_sparse(st["Vq_ptr"], st["Vq_idx"], st["Vq_val"], N)
    return jaccard_dense_from_sparse(jp, jj, jv, N)
