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
    print("N=%d D=%d k=%d: idx equal %s, keys equal %s, uncertified %d/%d, splits %d keep %d, max_err %.3e, exact %.3fs tc %.3fs"
          % (N, D, k, same, samek, info["uncertified_rows"], N, info["n_splits"], info["keep"],
             float(info["max_abs_err"]), te, tt), flush=True)
    if not same:
        bad = torch.nonzero((ie != it).any(dim=1)).flatten()
        print("  first bad rows", bad[:10].tolist())
        r = int(bad[0]); print("  exact", ie[r].tolist()); print("  tc   ", it[r].tolist())
    return same

ok = True
for cfg in [(512, 64, 16, 10), (2048, 256, 64, 30), (5000, 512, 5000, 30), (12936, 2048, 751, 30)]:
    ok &= run(*cfg)
print("ALL OK" if ok else "MISMATCH")
