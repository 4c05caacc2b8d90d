import os, sys
import torch, torch.distributed as dist
sys.path.insert(0, ".")
import reid_gan_b200 as rg
from reid_gan_b200 import sharded, knn_tc, faiss_rerank as fr
rank = int(os.environ["RANK"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
N = 32621
x, _ = rg.synth(N, 2048, 1041, 0.8, 0)
xd = x.cuda()
i1, k1_, info1 = fr.knn_search(xd, 30, "tc")
c1 = info1["cand_cnt"]
g_idx, g_key, info = sharded.knn_search_tiles(xd, 30)
W = dist.get_world_size(); b0, b1, B = sharded.block_partition(N, W, rank)
rc = info["cand_cnt"].view(W, B)
tot = rc.sum(0)[: b1 - b0]
print("rank", rank, "uncert", info["uncertified_rows"], "partial max", int(rc.max()), "sum-mismatch rows", int((tot != c1[b0:b1]).sum()),
      "equal idx", bool(torch.equal(g_idx, i1)), flush=True)
bad = torch.nonzero(tot != c1[b0:b1]).flatten()[:5]
for b in bad.tolist():
    print("  rank", rank, "row", b0 + b, "partials", rc[:, b].tolist(), "single", int(c1[b0 + b]), flush=True)
dist.destroy_process_group()
