"""Per-step device timing of the row-sharded pass (run under torchrun; rank 0 prints).  Every step is bracketed by
CUDA events on the launching stream inside an otherwise normal eager pass (so host launch gaps are included)."""
import os, sys, time
import torch, torch.distributed as dist
sys.path.insert(0, ".")
import reid_gan_b200 as rg
from reid_gan_b200 import _lib, sharded, pipeline

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
N = int(sys.argv[1]) if len(sys.argv) > 1 else 32621
x, _ = rg.synth(N, 2048, max(1, N // 31), 0.8, 0)
xd = x.cuda()
records = []
orig_call = _lib.call


def traced_call(name, *args):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); orig_call(name, *args); e1.record()
    records.append((name, e0, e1))


def wrap_coll(fn_name):
    fn = getattr(dist, fn_name)

    def w(*a, **k):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); r = fn(*a, **k); e1.record()
        nbytes = a[0].numel() * a[0].element_size()
        records.append(("nccl:%s[%.1fMB]" % (fn_name, nbytes / 1e6), e0, e1))
        return r
    setattr(dist, fn_name, w)


for plan in ("tiles", "tiles+rows"):
    for _ in range(5):
        sharded.pseudo_labels(xd, 30, 6, 0.6, 4, plan=plan)
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        sharded.pseudo_labels(xd, 30, 6, 0.6, 4, plan=plan)
    dist.barrier(); torch.cuda.synchronize()
    eager = (time.perf_counter() - t0) * 100
    for _ in range(3):
        sharded.pseudo_labels(xd, 30, 6, 0.6, 4, plan=plan, graph=True)
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        sharded.pseudo_labels(xd, 30, 6, 0.6, 4, plan=plan, graph=True)
    dist.barrier(); torch.cuda.synchronize()
    graph = (time.perf_counter() - t0) * 100
    # traced eager pass
    import reid_gan_b200.faiss_rerank as fr, reid_gan_b200.sharded as sh, reid_gan_b200.knn_tc as kt, reid_gan_b200.dbscan as db, reid_gan_b200.pipeline as pl
    mods = (fr, sh, kt, db, pl, _lib)
    saved = [(m, getattr(m, "call")) for m in mods if hasattr(m, "call")]
    for m, _ in saved:
        m.call = traced_call
    colls = ("all_gather_into_tensor", "all_to_all_single", "all_reduce")
    saved_c = [(c, getattr(dist, c)) for c in colls]
    for c in colls:
        wrap_coll(c)
    records.clear()
    es = torch.cuda.Event(enable_timing=True); ee = torch.cuda.Event(enable_timing=True)
    dist.barrier(); torch.cuda.synchronize()
    es.record()
    sharded.pseudo_labels(xd, 30, 6, 0.6, 4, plan=plan)
    ee.record(); torch.cuda.synchronize()
    for m, f in saved:
        m.call = f
    for c, f in saved_c:
        setattr(dist, c, f)
    if rank == 0:
        tot = es.elapsed_time(ee)
        busy = sum(a.elapsed_time(b) for _, a, b in records)
        print("== N=%d world=%d plan=%s: eager %.3f ms  graph %.3f ms  (traced pass %.3f ms, inside calls %.3f ms)" % (N, world, plan, eager, graph, tot, busy), flush=True)
        agg = {}
        for nm, a, b in records:
            c, t = agg.get(nm, (0, 0.0)); agg[nm] = (c + 1, t + a.elapsed_time(b))
        for nm, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            print("   %-46s x%d %8.3f ms" % (nm, c, t), flush=True)
from reid_gan_b200.pipeline import PassGraph
PassGraph._cache.clear()
torch.cuda.synchronize()
dist.barrier()
if rank == 0:
    print("TRACE DONE", flush=True)
os._exit(0)
