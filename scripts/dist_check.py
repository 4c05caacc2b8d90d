"""Multi-GPU check (run under torchrun): the row-sharded pass must be byte-identical to the single-GPU pass."""
import os, sys, time
import torch, torch.distributed as dist
sys.path.insert(0, ".")
import reid_gan_b200 as rg
from reid_gan_b200 import pipeline, sharded

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
ok = True
import itertools
cfgs = [(3000, 256, 100, 20, 6, 0.6), (12936, 2048, 751, 30, 6, 0.6), (32621, 2048, 1041, 30, 6, 0.6)]
if os.environ.get("ONLY_BIG"): cfgs = cfgs[-1:]
for (N, D, n_ids, k1, k2, eps), plan in itertools.product(cfgs, ("rows", "tiles", "tiles+rows")):
    if plan != "rows" and N < 8192:
        continue
    x, _ = rg.synth(N, D, n_ids, 0.8, 0)
    xd = x.cuda()
    ref = pipeline.pseudo_labels(xd, k1, k2, eps, 4, centroids=True)
    r0, r1 = sharded.partition(N, world, rank)
    out = sharded.pseudo_labels(xd[r0:r1].contiguous(), k1, k2, eps, 4, centroids=True, N=N, plan=plan)
    same = (torch.equal(out["labels"], ref["labels"]) and torch.equal(out["state"].rank, ref["state"].rank)
            and torch.equal(out["state"].Q_ptr, ref["state"].Q_ptr)
            and torch.equal(out["state"].Q_val[:ref["state"].q_total], ref["state"].Q_val[:ref["state"].q_total])
            and torch.equal(out["centroids"], ref["centroids"]))
    t = torch.tensor([1 if same else 0], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MIN)
    # timing
    for _ in range(3): sharded.pseudo_labels(xd, k1, k2, eps, 4, plan=plan)
    dist.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5): sharded.pseudo_labels(xd, k1, k2, eps, 4, plan=plan)
    dist.barrier(); torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
    o = sharded.pseudo_labels(xd, k1, k2, eps, 4, timers=True, plan=plan)
    print("   [rank %d] uncertified %s steps %s" % (rank, o["state"].knn_info.get("uncertified_rows"), o["state"].knn_info.get("steps_ms")), flush=True)
    if rank == 0:
        print("   stages(ms):", {k: round(v * 1e3, 3) for k, v in o["state"].timings.items()}, o["state"].knn_info.get("steps_ms"), flush=True)
    if rank == 0:
        print("N=%d world=%d plan=%s identical=%s clusters=%d  %.2f ms/pass" % (N, world, plan, bool(t.item()), int(out["num_clusters"]), dt * 1e3), flush=True)
    ok &= bool(t.item())
if rank == 0:
    print("DIST OK" if ok else "DIST MISMATCH")
dist.destroy_process_group()
