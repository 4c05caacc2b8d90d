"""Per-stage device times of the single-GPU pass for a few shapes. Usage: python scripts/stage_times.py [N n_ids]..."""
import sys
sys.path.insert(0, ".")
import torch
import reid_gan_b200 as rg
from reid_gan_b200 import pipeline

shapes = [(12936, 751), (32621, 1041)]
a = [int(v) for v in sys.argv[1:]]
if a:
    shapes = list(zip(a[0::2], a[1::2]))
for N, n_ids in shapes:
    x, _ = rg.synth(N, 2048, n_ids, 0.8, 0)
    x = x.cuda()
    for _ in range(3):
        out = pipeline.pseudo_labels(x, 30, 6, 0.6, 4)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        out = pipeline.pseudo_labels(x, 30, 6, 0.6, 4)
    e1.record()
    torch.cuda.synchronize()
    o = pipeline.pseudo_labels(x, 30, 6, 0.6, 4, timers=True)
    lab = out["labels"]
    print("N=%d ids=%d: %.3f ms/pass clusters=%d noise=%d edges=%d" % (N, n_ids, e0.elapsed_time(e1) / 10, int(out["num_clusters"].item()),
          int((lab < 0).sum()), int(out["nbr_cnt"].sum())), {k: round(v * 1e3, 3) for k, v in o["state"].timings.items()}, flush=True)
