"""Randomised stress of the symmetric tensor-core search against the exact search (GPU box, ~1 minute)."""
import sys, itertools
sys.path.insert(0, ".")
import torch
import reid_gan_b200 as rg
from reid_gan_b200 import faiss_rerank as fr

torch.manual_seed(0)
bad = 0
cases = []
for N in (8192, 8193, 8447, 9000, 12345, 16384, 20001):
    for D in (64, 192):
        for (n_ids, noise) in ((max(2, N // 31), 0.8), (N, 1.0), (8, 0.3)):
            cases.append((N, D, n_ids, noise))
for ci, (N, D, n_ids, noise) in enumerate(cases):
    k = (1, 5, 30, 32)[ci % 4]
    x, _ = rg.synth(N, D, n_ids, noise, ci)
    if ci % 3 == 0:
        x[N // 2: N // 2 + 300] = x[:300]             # duplicates
    if ci % 5 == 0:
        x = x * (0.5 + torch.rand(N, 1))               # rows that are not unit norm
    xd = x.cuda()
    ie, ke, _ = fr.knn_search(xd, k, "exact")
    it, kt, info = fr.knn_search(xd, k, "tc")
    ok = bool(torch.equal(ie, it)) and bool(torch.equal(ke, kt))
    bad += not ok
    print("N=%d D=%d ids=%d noise=%.1f k=%d mode=%s uncert=%d cand[min,max]=[%d,%d] %s" % (
        N, D, n_ids, noise, k, info["mode"], info["uncertified_rows"], int(info["cand_cnt"].min()), int(info["cand_cnt"].max()),
        "ok" if ok else "MISMATCH"), flush=True)
print("STRESS OK" if not bad else "STRESS FAILED: %d" % bad)
