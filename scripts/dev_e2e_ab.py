"""A/B of the chunked upload schedule of the host-features path (developer script): wall time of
compute_jaccard_distance + DBSCAN.fit_predict from pinned host features, uniform vs shrinking chunks, interleaved."""
import sys, time
sys.path.insert(0, ".")
import torch
import reid_gan_b200 as rg
from reid_gan_b200 import knn_tc as kt

N = int(sys.argv[1]) if len(sys.argv) > 1 else 32621
x = rg.synth(N, 2048, max(2, N // 31), 0.8, 0)[0].pin_memory()
uniform = kt.upload_bounds


def shrink(n_t, chunks):
    """last three chunks x 0.6 each"""
    w = [1.0] * max(1, chunks - 3) + [0.6, 0.36, 0.216][: max(0, min(3, chunks - 1))]
    acc, out = 0.0, [0]
    for v in w:
        acc += v
        out.append(int(round(n_t * acc / sum(w))))
    out[-1] = n_t
    return out


def run():
    d = rg.compute_jaccard_distance(x, k1=30, k2=6, print_flag=False, search_option=3)
    return rg.DBSCAN(eps=0.6, min_samples=4, metric="precomputed").fit_predict(d)


splits_on = kt.prepass_splits
counts = [int(c) for c in sys.argv[2].split(",")] if len(sys.argv) > 2 else [12, 16, 24]
variants = {"uniform%d" % c: (uniform, c, True) for c in counts}
variants["uniform%d-nosplit" % counts[0]] = (uniform, counts[0], False)
variants["shrink%d" % counts[0]] = (shrink, counts[0], True)
res = {k: [] for k in variants}
for rep in range(4):
    for name, (fn, ch, sp) in variants.items():
        kt.upload_bounds = fn
        kt.prepass_splits = splits_on if sp else (lambda n_rows, m, pair_slots=74: 1)
        kt.UPLOAD_CHUNKS = ch
        kt._tile_cache.clear()
        for _ in range(2):
            run()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(10):
            run()
        torch.cuda.synchronize()
        res[name].append((time.perf_counter() - t0) / 10 * 1e3)
for k, v in res.items():
    print(k, " ".join("%.3f" % t for t in v), "min %.3f" % min(v))
