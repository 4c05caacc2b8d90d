"""ORACLE support (test infrastructure): full-size parity fixtures.

    python -m oracle.make_fullsize [c1 c2 hard]

Runs the UNMODIFIED reference (clustercontrast/utils/faiss_rerank.py:30 compute_jaccard_distance with
search_option=3, through oracle/ref_shim.py) and sklearn's DBSCAN (examples/cluster_contrast_train_usl.py:160-163)
on the benchmark-size synthetic sets in the build container -- the dense path: three N x N float32 matrices,
13-17 GB at N = 32,621 -- and stores DIGESTS of the result under tests/golden/full_<name>.npz:

  rank_sha256              sha256 of initial_rank (int32, row-major)      faiss_rerank.py:62
  q_cnt, q_idx_sha256      nnz of every V_qe row / sha256 of the column indices (int32, row-major ascending)   :89-100
  nbr_cnt, nbr_idx_sha256  |{j : J_ij <= float32(eps)}| per row / sha256 of the ascending neighbour ids        :102-123 + sklearn
  labels                   DBSCAN(eps, min_samples=4, 'precomputed').fit_predict(J)  (int32)
  admissible               labels identical at eps - 1e-5 / eps / eps + 1e-5 (hard part 8)
  j_rows, j_vals           a few complete rows of J (float32)
  centroids_sha_rows/vals  rows 0, C//2, C-1 of the normalised centroids (train_usl.py:169-182,191)

The inputs are not stored: `reid_gan_b200.synth.synth / synth_hard` regenerate them from the seed (host RNG, same
bytes on every box; `x_sha256` guards that).  While the reference result is in memory the sparse restatement
(oracle/rerank.py sparse_pipeline + jaccard_sparse) is checked against it at full size, which is what pins the
oracle at N = 32,621 (tests/test_oracle.py re-checks the oracle against the digests).
"""
import hashlib
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shim, cluster, rerank as orr  # noqa: E402
from oracle.make_golden import _synth_mod  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")

# name -> (generator, kwargs, k1, k2, eps, min_samples)
CASES = {
    "c1": ("synth", dict(N=12936, D=2048, n_ids=751, noise=0.8, seed=0), 30, 6, 0.6, 4),      # BASELINE configs[0]
    "c2": ("synth", dict(N=32621, D=2048, n_ids=1041, noise=0.8, seed=0), 30, 6, 0.6, 4),     # BASELINE configs[1]
    "hard": ("synth_hard", dict(N=20480, D=2048, seed=0), 30, 6, 0.6, 4),
}


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def make_inputs(name, sm=None):
    sm = sm or _synth_mod()
    gen, kw, k1, k2, eps, ms = CASES[name]
    x, ids = getattr(sm, gen)(**kw)
    return x, ids, k1, k2, eps, ms


def digest_from_dense(J, rank, x, eps, ms):
    """Digest dict from the reference's dense J (N x N float32) and initial_rank."""
    N = J.shape[0]
    eps32 = np.float32(eps)
    out = {}
    out["rank_sha256"] = sha(rank.astype(np.int32))
    nbr_cnt = np.empty(N, np.int32)
    h = hashlib.sha256()
    for a in range(0, N, 2048):
        m = J[a:a + 2048] <= eps32
        nbr_cnt[a:a + 2048] = m.sum(1)
        h.update(np.nonzero(m)[1].astype(np.int32).tobytes())
    out["nbr_cnt"] = nbr_cnt
    out["nbr_idx_sha256"] = h.hexdigest()
    lab = cluster.sklearn_dbscan(J, eps, ms)
    lo = cluster.sklearn_dbscan(J, eps - 1e-5, ms)
    hi = cluster.sklearn_dbscan(J, eps + 1e-5, ms)
    out["labels"] = lab.astype(np.int32)
    out["admissible"] = bool(np.array_equal(lab, lo) and np.array_equal(lab, hi))
    rows = np.unique(np.concatenate([np.arange(0, N, max(1, N // 12))[:12], [N - 1]]))
    out["j_rows"] = rows.astype(np.int32)
    out["j_vals"] = J[rows].copy()
    if lab.max() >= 0:
        cen = ref_shim.ref_generate_cluster_features(lab, torch.from_numpy(x)).numpy()
        C = cen.shape[0]
        cr = np.unique([0, C // 2, C - 1])
        out["centroid_rows"] = cr.astype(np.int32)
        out["centroid_vals"] = cen[cr]
        out["num_clusters"] = C
    return out


def run_case(name):
    x_t, ids, k1, k2, eps, ms = make_inputs(name)
    x = x_t.numpy()
    N = x.shape[0]
    t0 = time.perf_counter()
    mod = ref_shim.load_faiss_rerank()
    torch.set_num_threads(1)                      # the per-row torch.mm of faiss_rerank.py:83 thrashes with more
    J = mod.compute_jaccard_distance(x_t, k1=k1, k2=k2, print_flag=False, search_option=3)
    t_ref = time.perf_counter() - t0
    assert J.dtype == np.float32 and J.shape == (N, N)
    rank = orr.exact_knn(x, k1)                   # what the faiss stand-in returned inside the reference call
    t0 = time.perf_counter()
    out = digest_from_dense(J, rank, x, eps, ms)
    t_db = time.perf_counter() - t0
    # ---- pin the sparse restatement at this size ------------------------------------------------
    st = orr.sparse_pipeline(x, k1, k2, rank=rank)
    qp, qi, qv = st["Vq_ptr"], st["Vq_idx"], st["Vq_val"]
    jp, jj, jv = orr.jaccard_sparse(qp, qi, qv, N)
    rows = np.repeat(np.arange(N), np.diff(jp))
    # the reference's structure: J != 1 exactly on the pairs that share a column (clip-to-0 pairs included)
    n_ref_pairs = 0
    worst = 0.0
    for a in range(0, N, 2048):
        b = min(N, a + 2048)
        sel = (rows >= a) & (rows < b)
        blk = np.ones((b - a, N), np.float32)
        blk[rows[sel] - a, jj[sel]] = jv[sel]
        ref = J[a:b]
        worst = max(worst, float(np.abs(blk - ref).max()))
        n_ref_pairs += int((ref != 1.0).sum())
        # every pair the reference moved off 1.0 must be a stored pair
        assert not ((ref != 1.0) & (blk == 1.0)).any(), "reference has a pair the sparse oracle lacks"
    assert worst <= 1e-6, "sparse oracle J differs from the reference by %g" % worst
    out["q_cnt"] = np.diff(qp).astype(np.int32)
    out["q_idx_sha256"] = sha(qi.astype(np.int32))
    out["e_cnt"] = np.diff(st["E_ptr"]).astype(np.int32)
    out["e_idx_sha256"] = sha(st["E_idx"].astype(np.int32))
    out["x_sha256"] = sha(x)
    out["oracle_vs_reference_max_abs_J"] = worst
    out["reference_seconds"] = t_ref
    out["dbscan_digest_seconds"] = t_db
    out["k1"], out["k2"], out["eps"], out["min_samples"], out["N"] = k1, k2, eps, ms, N
    np.savez_compressed(os.path.join(GOLD, "full_%s.npz" % name), **out)
    lab = out["labels"]
    print("full", name, "N", N, "reference %.1f s" % t_ref, "clusters", int(lab.max()) + 1, "noise", int((lab < 0).sum()),
          "admissible", out["admissible"], "oracle max|dJ| %.2e" % worst, "edges", int(out["nbr_cnt"].sum()), flush=True)


if __name__ == "__main__":
    if not ref_shim.available():
        raise SystemExit("the reference tree is not present: fixtures can only be generated in the build container")
    for nm in (sys.argv[1:] or list(CASES)):
        run_case(nm)
