"""ORACLE support: generate tests/golden/*.npz by running the UNMODIFIED reference
Python (via oracle/ref_shim.py) in the build container.  Re-run with

    python -m oracle.make_golden

Every fixture stores its inputs next to the reference outputs, so the tests never
need /root/reference.  Reference entry points exercised:
  clustercontrast/utils/faiss_rerank.py:30   compute_jaccard_distance(search_option=3)
  sklearn.cluster.DBSCAN(...).fit_predict    (examples/cluster_contrast_train_usl.py:160,163)
  examples/cluster_contrast_train_usl.py:169-182,191  (closure twin in ref_shim)
  clustercontrast/models/cm.py:36,75,125-135 cm / cm_hard + ClusterMemory.forward body (CPU)
  clustercontrast/utils/infomap_cluster.py:230-234, 129-144   get_dist_nbr / get_links            (next row f1)
  clustercontrast/utils/rerank.py:31-97                       re_ranking                          (next row f2)
  clustercontrast/evaluation_metrics/ranking.py:18-115 + evaluators.py:78-87   cmc / mean_ap / pairwise distance (f3)
"""
import importlib.util
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shim, cluster  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def _synth_mod():
    spec = importlib.util.spec_from_file_location("_synth", os.path.join(ROOT, "reid-gan_b200", "synth.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


RERANK_CASES = [
    # name,            N,    D,  n_ids, noise, seed, k1, k2, dup
    ("n512_k20",       512,  64,  16,   0.8,   0,    20, 6,  0),
    ("n1024_k30",      1024, 128, 33,   0.8,   1,    30, 6,  0),
    ("n768_k15_k2_1",  768,  64,  24,   0.8,   2,    15, 1,  0),   # odd k1 (half-to-even), k2 == 1
    ("n640_k25_dups",  640,  64,  20,   0.8,   3,    25, 4,  40),  # duplicate rows: self not at rank 0
    ("n500_k30_noisy", 500,  64,  170,  1.2,   4,    30, 6,  0),   # many noise points for DBSCAN
]
EPS_LIST = (0.4, 0.5, 0.6)


def make_rerank(sm):
    for name, N, D, n_ids, noise, seed, k1, k2, dup in RERANK_CASES:
        x, ids = sm.synth(N, D, n_ids, noise, seed)
        if dup:
            g = torch.Generator().manual_seed(seed)
            src = torch.randint(0, N, (dup,), generator=g)
            dst = torch.randint(0, N, (dup,), generator=g)
            x[dst] = x[src]
        J = ref_shim.ref_compute_jaccard_distance(x, k1, k2)
        assert J.dtype == np.float32 and J.shape == (N, N)
        rows, cols = np.nonzero(J != 1.0)
        out = dict(x=x.numpy(), k1=k1, k2=k2, J_rows=rows.astype(np.int32), J_cols=cols.astype(np.int32),
                   J_vals=J[rows, cols], eps_list=np.asarray(EPS_LIST))
        for eps in EPS_LIST:
            lab = cluster.sklearn_dbscan(J, eps, 4)
            out["labels_eps%02d" % round(eps * 100)] = lab.astype(np.int64)
            # admissibility band (SURVEY section 7 hard part 8)
            lo = cluster.sklearn_dbscan(J, eps - 1e-5, 4)
            hi = cluster.sklearn_dbscan(J, eps + 1e-5, 4)
            out["admissible_eps%02d" % round(eps * 100)] = bool(np.array_equal(lab, lo) and np.array_equal(lab, hi))
            if lab.max() >= 0:
                cen = ref_shim.ref_generate_cluster_features(lab, x)
                out["centroids_eps%02d" % round(eps * 100)] = cen.numpy()
        np.savez_compressed(os.path.join(GOLD, "rerank_%s.npz" % name), **out)
        print("rerank", name, "pairs<1:", rows.size, {k: v for k, v in out.items() if k.startswith("admissible")})


def make_rerank_half(sm):
    """use_float16=True (faiss_rerank.py:37): float16 V / V_qe / sums / distances."""
    for name, N, D, n_ids, noise, seed, k1, k2 in [("n400_k20", 400, 64, 14, 0.8, 7, 20, 6),
                                                   ("n300_k15_k2_1", 300, 64, 10, 0.8, 8, 15, 1)]:
        x, _ = sm.synth(N, D, n_ids, noise, seed)
        torch.set_num_threads(1)
        mod = ref_shim.load_faiss_rerank()
        J = mod.compute_jaccard_distance(x, k1=k1, k2=k2, print_flag=False, search_option=3, use_float16=True)
        assert J.dtype == np.float16 and J.shape == (N, N)
        rows, cols = np.nonzero(J != 1.0)
        np.savez_compressed(os.path.join(GOLD, "halfrerank_%s.npz" % name), x=x.numpy(), k1=k1, k2=k2, use_float16=True,
                            J_rows=rows.astype(np.int32), J_cols=cols.astype(np.int32), J_vals=J[rows, cols])
        print("rerank", name, "pairs<1:", rows.size, J.dtype)


def make_cm(sm):
    cm_mod = ref_shim.load_cm()
    torch.set_num_threads(1)
    for name, C, D, n_lab, n_inst, temp, mom, seed in [
        ("c40_d64", 40, 64, 8, 4, 0.05, 0.2, 0),
        ("c300_d128", 300, 128, 16, 16, 0.05, 0.2, 1),
        ("c33_d128_single", 33, 128, 32, 1, 0.07, 0.1, 2),     # every label present once
    ]:
        g = torch.Generator().manual_seed(seed)
        feats = F.normalize(torch.randn(C, D, generator=g), dim=1)
        labs = torch.randperm(C, generator=g)[:n_lab]
        targets = labs.repeat_interleave(n_inst)[torch.randperm(n_lab * n_inst, generator=g)]
        inputs = feats[targets] + 2.5 * torch.randn(targets.numel(), D, generator=g) / D ** 0.5
        inputs = inputs * (1.0 + 0.1 * torch.randn(targets.numel(), 1, generator=g)).abs()
        gl = torch.rand(targets.numel(), generator=g) + 0.5
        out = dict(inputs=inputs.numpy(), targets=targets.numpy(), features=feats.numpy(),
                   temp=temp, momentum=mom, grad_loss=gl.numpy())
        for hard in (False, True):
            f = feats.clone()
            x = inputs.clone().requires_grad_(True)
            xn = F.normalize(x, dim=1)                       # cm.py:125 (without .cuda())
            logits = (cm_mod.cm_hard if hard else cm_mod.cm)(xn, targets, f, mom)   # :128/:130
            logits = logits / temp                           # :134 (out of place: same values)
            loss = F.cross_entropy(logits, targets, reduction="none")               # :135
            loss.backward(gl)
            tag = "hard" if hard else "cm"
            out["loss_" + tag] = loss.detach().numpy()
            out["grad_" + tag] = x.grad.numpy()
            out["features_after_" + tag] = f.numpy()
        np.savez_compressed(os.path.join(GOLD, "cm_%s.npz" % name), **out)
        print("cm", name, "loss mean", float(out["loss_cm"].mean()))


def _pdist(a, b):
    """evaluators.py:78-87 pairwise_distance (the reference function takes feature dicts; this is its arithmetic)."""
    m, n = a.size(0), b.size(0)
    d = torch.pow(a, 2).sum(dim=1, keepdim=True).expand(m, n) + torch.pow(b, 2).sum(dim=1, keepdim=True).expand(n, m).t()
    d = d.clone()
    d.addmm_(a, b.t(), beta=1, alpha=-2)
    return d


def make_next_rows(sm):
    """f1 / f2 / f3: outputs of the unmodified reference functions on small seeded inputs."""
    torch.set_num_threads(1)
    # f1 -- infomap front end
    im = ref_shim.load_infomap_cluster()
    x, _ = sm.synth(400, 64, 16, 0.8, 11)
    dists, nbrs = im.get_dist_nbr(features=x.numpy(), k=15, knn_method='faiss-cpu')
    out = dict(x=x.numpy(), k=15, dists=dists, nbrs=nbrs)
    for min_sim in (0.3, 0.5):
        single, links = im.get_links(single=[], links={}, nbrs=nbrs, dists=dists, min_sim=min_sim)
        tag = "ms%02d" % round(min_sim * 100)
        keys = np.array(sorted(links.keys()), dtype=np.int64).reshape(-1, 2)
        out["links_ij_" + tag] = keys
        out["links_w_" + tag] = np.array([links[(int(a), int(b))] for a, b in keys], dtype=np.float64)
        out["single_" + tag] = np.array(single, dtype=np.int64)
    np.savez_compressed(os.path.join(GOLD, "infomap_n400_k15.npz"), **out)
    print("infomap", len(out["links_w_ms50"]), "links at 0.5")
    # f2 -- evaluation re-ranking
    rr = ref_shim.load_eval_rerank()
    x, _ = sm.synth(360, 64, 24, 0.8, 12)
    q, g = x[:100], x[100:]
    qg, qq, gg = _pdist(q, g).numpy(), _pdist(q, q).numpy(), _pdist(g, g).numpy()
    out = dict(q_g=qg, q_q=qq, g_g=gg)
    for k1, k2, lam in ((20, 6, 0.3), (7, 1, 0.5)):
        out["final_k%d_%d" % (k1, k2)] = rr.re_ranking(qg, qq, gg, k1=k1, k2=k2, lambda_value=lam)
        out["lambda_k%d_%d" % (k1, k2)] = lam
    np.savez_compressed(os.path.join(GOLD, "evalrerank_q100_g260.npz"), **out)
    print("eval rerank", out["final_k20_6"].shape)
    # f3 -- ranking metrics
    rk = ref_shim.load_ranking()
    x, ids = sm.synth(500, 32, 25, 1.2, 13)
    cams = np.random.default_rng(13).integers(0, 4, 500)
    q, g = x[:120], x[120:]
    dm = _pdist(q, g)
    qi, gi, qc, gc = ids[:120].numpy(), ids[120:].numpy(), cams[:120], cams[120:]
    out = dict(q=q.numpy(), g=g.numpy(), distmat=dm.numpy(), q_ids=qi, g_ids=gi, q_cams=qc, g_cams=gc,
               mAP=np.float64(rk.mean_ap(dm, qi, gi, qc, gc)),
               cmc_market=rk.cmc(dm, qi, gi, qc, gc, topk=50, separate_camera_set=False, single_gallery_shot=False,
                                 first_match_break=True),
               cmc_allshots=rk.cmc(dm, qi, gi, qc, gc, topk=50, separate_camera_set=False, single_gallery_shot=False,
                                   first_match_break=False),
               cmc_sepcam=rk.cmc(dm, qi, gi, qc, gc, topk=50, separate_camera_set=True, single_gallery_shot=False,
                                 first_match_break=True))
    np.savez_compressed(os.path.join(GOLD, "ranking_q120_g380.npz"), **out)
    print("ranking mAP", float(out["mAP"]))
    # single_gallery_shot=True draws from the process-wide np.random state (ranking.py:10-16): golden per seed
    sgs = {}
    for name, kw in (("sgs", dict()), ("sgs_fmb", dict(first_match_break=True)), ("sgs_sep", dict(separate_camera_set=True))):
        np.random.seed(1234)
        sgs["cmc_" + name + "_seed1234"] = rk.cmc(dm, qi, gi, qc, gc, topk=50, single_gallery_shot=True, **kw)
    np.savez_compressed(os.path.join(GOLD, "ranking_sgs_q120_g380.npz"), **sgs)


if __name__ == "__main__":
    os.makedirs(GOLD, exist_ok=True)
    sm = _synth_mod()
    if "--next-only" not in sys.argv:
        make_rerank(sm)
        make_rerank_half(sm)
        make_cm(sm)
    make_next_rows(sm)
