"""One pseudo-label pass inside a cudaProfilerStart/Stop bracket, for `ncu --profile-from-start off`.
Usage: python scripts/prof_pass.py [N] [n_ids] [passes]"""
import sys
sys.path.insert(0, ".")
import torch
import reid_gan_b200 as rg
from reid_gan_b200 import pipeline

N = int(sys.argv[1]) if len(sys.argv) > 1 else 32621
n_ids = int(sys.argv[2]) if len(sys.argv) > 2 else 1041
passes = int(sys.argv[3]) if len(sys.argv) > 3 else 1
x, _ = rg.synth(N, 2048, n_ids, 0.8, 0)
x = x.cuda()
for _ in range(2):
    pipeline.pseudo_labels(x, 30, 6, 0.6, 4)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
for _ in range(passes):
    out = pipeline.pseudo_labels(x, 30, 6, 0.6, 4)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("clusters", int(out["num_clusters"].item()))
