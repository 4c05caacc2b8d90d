"""Small end-to-end run that launches every kernel family once (the whole public surface in ~5 s on a B200):
pseudo-label pass (symmetric and one-sided search, streamed upload), ClusterMemory, and the f1-f3 rows."""
import sys
sys.path.insert(0, ".")
import numpy as np
import torch
import reid_gan_b200 as rg
from reid_gan_b200 import pipeline, infomap_cluster as ic, evaluation as ev
from oracle import eval_rerank as oe

N, D = 8704, 128
x, ids = rg.synth(N, D, 280, 0.8, 0)
out = pipeline.pseudo_labels(x.cuda(), 30, 6, 0.6, 4, centroids=True)          # sym search, grouped re-score, sparse stages
print("pass ok", int(out["num_clusters"].item()), out["state"].knn_info["mode"])
d = rg.compute_jaccard_distance(x[:3000].contiguous(), k1=20, k2=6, print_flag=False)   # one-sided kernel path
lab = rg.DBSCAN(eps=0.6, min_samples=4).fit_predict(d)
J = d.dense_device(0, 64)
print("small ok", int(lab.max()) + 1, float(J.min()))
d2 = rg.compute_jaccard_distance(x.pin_memory(), k1=30, k2=6, print_flag=False)         # streamed upload
print("upload ok", d2.state.knn_info["sym"])
cen = rg.generate_cluster_features(out["labels"].cpu().numpy(), x, normalize=True)
mem = rg.ClusterMemory(D, cen.shape[0], temp=0.05, momentum=0.2, use_hard=True).cuda()
mem.features = cen.clone()
inp, tgt = rg.synth_cm_batch(x, None, out["labels"].cpu(), num_ids=8, num_instances=4, seed=0)
inp = inp.cuda().requires_grad_(True)
mem(inp, tgt.cuda()).mean().backward()
print("cm ok")
dd, nn_ = ic.get_dist_nbr(x[:2000].numpy(), k=15)
s, l = ic.get_links([], {}, nn_, dd, 0.5)
qg, qq, gg = oe.synthetic_distances(600, 150, 32, 30, 3)
r = rg.re_ranking(qg, qq, gg)
print("f1/f2 ok", len(l), r.shape)
m = ev.mean_ap(qg, np.arange(150) % 30, np.arange(450) % 30, np.zeros(150, int), np.ones(450, int))
c = ev.cmc(qg, np.arange(150) % 30, np.arange(450) % 30, np.zeros(150, int), np.ones(450, int), topk=10, first_match_break=True)
print("f3 ok", float(m), c[:3])
torch.cuda.synchronize()
