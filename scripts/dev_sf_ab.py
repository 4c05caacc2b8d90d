"""A/B of the symmetric search layouts (knn_tc.SYM_SAMPLE_FIRST): identical neighbour lists and keys, per-entry-point
CUDA-event times of the candidate stages.  Usage: python scripts/dev_sf_ab.py [N] [synth|hard]"""
import sys
sys.path.insert(0, ".")
import torch
import reid_gan_b200 as rg
from reid_gan_b200 import _lib, knn_tc as kt, faiss_rerank as fr

N = int(sys.argv[1]) if len(sys.argv) > 1 else 32621
gen = sys.argv[2] if len(sys.argv) > 2 else "synth"
D = 2048 if N > 10000 else 256
x = rg.synth(N, D, max(1, N // 31), 0.8, 0)[0] if gen == "synth" else rg.synth_hard(N, D, seed=0)[0]
x = x.cuda()
res = {}
for sf in (False, True, False, True):
    kt.SYM_SAMPLE_FIRST = sf
    for _ in range(2):
        idx, key, info = fr.knn_search(x, 30, "tc")
    torch.cuda.synchronize()
    _lib.profiler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(8):
        idx, key, info = fr.knn_search(x, 30, "tc")
    e1.record()
    torch.cuda.synchronize()
    _lib.profiler.stop()
    prof = {k: round(v[1] / 8, 4) for k, v in _lib.profiler.summary().items()}
    cnt = info["cand_cnt"]
    print("%s N=%d sample_first=%s: %.3f ms/search (eager, with event overhead); uncertified=%d; list mean %.1f min %d max %d; %s" % (
        gen, N, sf, e0.elapsed_time(e1) / 8, info["uncertified_rows"], float(cnt.float().mean()), int(cnt.min()), int(cnt.max()),
        info["sym"]), flush=True)
    print("   ", prof, flush=True)
    res[sf] = (idx.clone(), key.clone())
same = torch.equal(res[False][0], res[True][0]) and torch.equal(res[False][1], res[True][1])
print("identical neighbour lists and keys:", same, flush=True)
if N <= 40000:
    ie, ke, _ = fr.knn_search(x, 30, "exact")
    print("identical to the exact search:", torch.equal(ie, res[True][0]) and torch.equal(ke, res[True][1]), flush=True)
assert same
if len(sys.argv) > 3 and sys.argv[3] == "prof":           # per-kernel device times (CUPTI), both layouts
    from torch.profiler import profile, ProfilerActivity
    for sf in (False, True):
        kt.SYM_SAMPLE_FIRST = sf
        fr.knn_search(x, 30, "tc")
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(5):
                fr.knn_search(x, 30, "tc")
            torch.cuda.synchronize()
        print("sample_first=%s" % sf)
        for e in sorted(prof.key_averages(), key=lambda e: -e.device_time_total)[:14]:
            print("   %-60s n=%3d  %.1f us/pass" % (e.key[:60], e.count, e.device_time_total / 5))
