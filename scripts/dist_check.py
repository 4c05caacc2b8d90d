"""Multi-GPU check (run under torchrun): the row-sharded pass must be byte-identical to the single-GPU pass -- in the
sync-free flavour (eager and replayed from a CUDA graph), in the exact-size flavour, from pre-sharded features (f4),
and when the speculative sizes are forced to fail.  Prints one digest line per configuration.
    python -m torch.distributed.run --nproc-per-node W --master-addr 127.0.0.1 scripts/dist_check.py"""
import hashlib, os, sys, time, itertools
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, ".")
import reid_gan_b200 as rg
from reid_gan_b200 import pipeline, sharded, faiss_rerank as fr

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
ok = True


def sha(t):
    return hashlib.sha256(np.ascontiguousarray(t.cpu().numpy()).tobytes()).hexdigest()[:16]


def same_as(ref, out):
    nq = ref["state"].q_total
    return (torch.equal(out["labels"], ref["labels"]) and torch.equal(out["state"].rank, ref["state"].rank)
            and torch.equal(out["state"].Q_ptr, ref["state"].Q_ptr)
            and torch.equal(out["state"].Q_idx[:nq], ref["state"].Q_idx[:nq])
            and torch.equal(out["state"].Q_val[:nq], ref["state"].Q_val[:nq])
            and ("centroids" not in ref or torch.equal(out["centroids"], ref["centroids"])))


def agree(flag):
    t = torch.tensor([1 if flag else 0], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return bool(t.item())


def timed(fn, n=10):
    for _ in range(3): fn()
    dist.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    dist.barrier(); torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


cfgs = [("synth", dict(N=3000, D=256, n_ids=100, noise=0.8, seed=0), 20, 6, 0.6),
        ("synth", dict(N=12936, D=2048, n_ids=751, noise=0.8, seed=0), 30, 6, 0.6),
        ("synth_hard", dict(N=20480, D=2048, seed=0), 30, 6, 0.6),
        ("synth", dict(N=32621, D=2048, n_ids=1041, noise=0.8, seed=0), 30, 6, 0.6)]
if os.environ.get("ONLY_BIG"): cfgs = cfgs[-1:]
for (gen, kw, k1, k2, eps), plan in itertools.product(cfgs, ("rows", "tiles", "tiles+rows")):
    N = kw["N"]
    if plan != "rows" and N < 8192:
        continue
    if plan == "rows" and N > 13000:
        continue
    x, _ = getattr(rg, gen)(**kw)
    xd = x.cuda()
    ref = pipeline.pseudo_labels(xd, k1, k2, eps, 4, centroids=True)
    r0, r1 = sharded.partition(N, world, rank)
    x_local = xd[r0:r1].contiguous()                      # f4: every rank starts from ITS rows only
    res = {}
    out = sharded.pseudo_labels(x_local, k1, k2, eps, 4, centroids=True, N=N, plan=plan)
    res["sync-free, pre-sharded rows"] = same_as(ref, out)
    res["sync-free, replicated rows"] = same_as(ref, sharded.pseudo_labels(xd, k1, k2, eps, 4, centroids=True, plan=plan))
    res["exact sizes"] = same_as(ref, sharded.pseudo_labels(xd, k1, k2, eps, 4, centroids=True, plan=plan, speculative=False))
    if plan != "rows":
        o = sharded.pseudo_labels(xd, k1, k2, eps, 4, centroids=True, plan=plan, graph=True)
        o = sharded.pseudo_labels(xd, k1, k2, eps, 4, centroids=True, plan=plan, graph=True)
        res["graph replay"] = same_as(ref, o)
        keep = dict(fr.REC_STRIDE), fr.QE_SPEC_SLOTS
        fr.REC_STRIDE.update(V=8, Q=8, nbr=8); fr._rec_stride_hint.clear()
        o = sharded.pseudo_labels(xd, k1, k2, eps, 4, centroids=True, plan=plan)
        res["forced record overflow -> redo"] = same_as(ref, o) and ("speculation_failed" in o["state"].knn_info or plan == "tiles")
        fr.REC_STRIDE.update(keep[0]); fr._rec_stride_hint.clear()
    good = all(agree(v) for v in res.values())
    ok &= good
    t_e = timed(lambda: sharded.pseudo_labels(xd, k1, k2, eps, 4, plan=plan))
    t_g = timed(lambda: sharded.pseudo_labels(xd, k1, k2, eps, 4, plan=plan, graph=True)) if plan != "rows" else float("nan")
    if os.environ.get("REID_TRACE_STEPS", "0") != "0":             # this script's own switch; the package reads no environment
        sharded.TRACE_STEPS = True
        o = sharded.pseudo_labels(xd, k1, k2, eps, 4, plan=plan)
        sharded.TRACE_STEPS = False
        print("   [rank %d] steps %s" % (rank, o["state"].knn_info.get("steps_ms")), flush=True)
    if rank == 0:
        print("%s N=%d world=%d plan=%-10s identical=%s %s labels=%s rank=%s clusters=%d noise=%d  eager %.2f ms  graph %.2f ms"
              % (gen, N, world, plan, good, {k: v for k, v in res.items() if not v} or "", sha(out["labels"]), sha(out["state"].rank),
                 int(out["num_clusters"]), int((out["labels"] < 0).sum()), t_e, t_g), flush=True)
if rank == 0:
    print("DIST OK" if ok else "DIST MISMATCH", flush=True)
# graphs that hold NCCL kernels must go before the communicator does (a destroy with live graphs hangs)
pipeline.PassGraph._cache.clear()
torch.cuda.synchronize()
dist.barrier()
os._exit(0)
