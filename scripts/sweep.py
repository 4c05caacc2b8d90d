"""Scale sweep (BASELINE configs[4]): full pseudo-label pass at N = 100k / 250k synthetic features x 2048-d,
row/tile-partitioned over the ranks of one box (or a single GPU).  Checks size-independent properties and, when
asked (CHECK_SINGLE=1), byte-identity with the single-GPU pass.
    python scripts/sweep.py 100000 [250000]            (1 GPU)
    torchrun --nproc-per-node 8 scripts/sweep.py ...   (W GPUs)"""
import json, os, sys, time
sys.path.insert(0, ".")
import torch
import torch.distributed as dist
import reid_gan_b200 as rg
from reid_gan_b200 import pipeline, sharded

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); lr = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
sizes = [int(v) for v in sys.argv[1:]] or [100000]
steps = int(os.environ.get("STEPS", "3"))
for N in sizes:
    n_ids = max(1, round(N * 1041 / 32621))          # the workload's cluster size
    x, ids = rg.synth_device(N, 2048, n_ids, 0.8, 0, "cuda")
    run = (lambda: sharded.pseudo_labels(x, 30, 6, 0.6, 4)) if world > 1 else (lambda: pipeline.pseudo_labels(x, 30, 6, 0.6, 4))
    out = run(); out = run()
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        out = run()
    e1.record()
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    if world > 1:
        t = torch.tensor([ms], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t)
    lab = out["labels"]; st = out["state"]
    # properties: self is the nearest neighbour; each pseudo-label is pure w.r.t. the generating identity; the
    # number of clusters equals the number of identities that have >= min_samples members
    self_first = bool((st.rank[:, 0].long() == torch.arange(N, device="cuda")).all())
    ncl = int(out["num_clusters"].item())
    ok_lab = lab >= 0
    pure = True
    if ncl:
        first_id = torch.full((ncl,), -1, dtype=torch.int64, device="cuda")
        first_id[lab[ok_lab]] = ids[ok_lab]
        pure = bool((first_id[lab[ok_lab]] == ids[ok_lab]).all())
    same = None
    if os.environ.get("CHECK_SINGLE") and world > 1:
        ref = pipeline.pseudo_labels(x, 30, 6, 0.6, 4)
        same = bool(torch.equal(ref["labels"], lab) and torch.equal(ref["state"].rank, st.rank))
    if rank == 0:
        info = st.knn_info
        print(json.dumps({"N": N, "n_ids": n_ids, "gpus": world, "ms_per_pass": round(ms, 3), "clusters": ncl,
                          "noise_points": int((lab < 0).sum()), "self_first": self_first, "labels_pure": pure,
                          "identical_to_single_gpu": same, "knn": info.get("mode"), "uncertified_rows": info.get("uncertified_rows"),
                          "peak_mem_gb": round(torch.cuda.max_memory_allocated() / 2**30, 2)}), flush=True)
    del out, x
    torch.cuda.empty_cache()
if world > 1:
    dist.destroy_process_group()
