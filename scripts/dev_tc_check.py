"""Developer check for the tcgen05 candidate kernel (run on the GPU box under `timeout`)."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import reid_gan_b200 as rg
from reid_gan_b200 import faiss_rerank as fr

def run(N, D, n_ids, k, noise=0.8, seed=0):
    x, _ = rg.synth(N, D, n_ids, noise, seed)
    xd = x.cuda()
    torch.cuda.synchronize()
    t = time.time(); ie, ke, _ = fr.knn_search(xd, k, "exact"); torch.cuda.synchronize(); te = time.time() - t
    t = time.time(); it, kt, info = fr.knn_search(xd, k, "tc"); torch.cuda.synchronize(); tt = time.time() - t
    same = bool(torch.equal(ie, it)); samek = bool(torch.equal(ke, kt))
    extra = ""
    if info.get("cand_cnt") is not None:
        c = info["cand_cnt"].float()
        extra = " | sym %s cand/row mean %.0f min %d max %d" % (info["sym"], c.mean(), int(c.min()), int(c.max()))
    print("N=%d D=%d k=%d [%s]: idx equal %s, keys equal %s, uncertified %d/%d, splits %d keep %d, max_err %.3e, exact %.3fs tc %.3fs%s"
          % (N, D, k, info["mode"], same, samek, info["uncertified_rows"], N, info["n_splits"], info["keep"],
             float(info["max_abs_err"]), te, tt, extra), flush=True)
    if not same:
        bad = torch.nonzero((ie != it).any(dim=1)).flatten()
        print("  first bad rows", bad[:10].tolist())
        r = int(bad[0]); print("  exact", ie[r].tolist()); print("  tc   ", it[r].tolist())
    return same

ok = True
cfgs = [(512, 64, 16, 10), (2048, 256, 64, 30), (5000, 512, 5000, 30), (12936, 2048, 751, 30), (9000, 512, 9000, 30),
        (20000, 256, 100, 30), (32621, 2048, 1041, 30)]
for cfg in cfgs:
    ok &= run(*cfg)
print("ALL OK" if ok else "MISMATCH")
