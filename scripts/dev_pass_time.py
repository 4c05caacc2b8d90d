"""Developer timing: wall-clock per pass, eager vs graph replay (one GPU)."""
import sys, time
sys.path.insert(0, ".")
import torch
import reid_gan_b200 as rg
from reid_gan_b200 import pipeline, _lib

N = int(sys.argv[1]) if len(sys.argv) > 1 else 32621
x, _ = rg.synth(N, 2048, max(1, N // 31), 0.8, 0)
x = x.cuda()
for mode in ("eager", "graph", "eager"):
    ts = []
    for i in range(12):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = pipeline.pseudo_labels(x, 30, 6, 0.6, 4, graph=(mode == "graph"))
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t0) * 1e3)
    print(mode, " ".join("%.2f" % t for t in ts), "alloc MB", torch.cuda.memory_allocated() >> 20, "reserved MB", torch.cuda.memory_reserved() >> 20, flush=True)
# back-to-back (no sync between passes), like bench.py
for mode in ("eager", "graph"):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(20):
        out = pipeline.pseudo_labels(x, 30, 6, 0.6, 4, graph=(mode == "graph"))
    torch.cuda.synchronize()
    print(mode, "back-to-back %.3f ms/pass" % ((time.perf_counter() - t0) * 1e3 / 20), flush=True)
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for i in range(5):
    out = pipeline.pseudo_labels(x, 30, 6, 0.6, 4)
torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
