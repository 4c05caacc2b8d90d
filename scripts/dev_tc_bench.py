"""Time the tcgen05 candidate kernel alone (developer tool): python scripts/dev_tc_bench.py [N] [reps]"""
import ctypes, sys, os
import torch
sys.path.insert(0, ".")
import reid_gan_b200 as rg
from reid_gan_b200 import _lib, knn_tc
from reid_gan_b200._lib import call, ptr, stream_ptr

N = int(sys.argv[1]) if len(sys.argv) > 1 else 32621
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
D, k = 2048, 30
x, _ = rg.synth(N, D, max(1, N // 31), 0.8, 0)
x = x.cuda()
xh = torch.empty((N, D), dtype=torch.float16, device="cuda")
call("reid_features_to_half", ptr(x), N, D, knn_tc.SCALE_LOG2, ptr(xh), None, stream_ptr())
for cg in ([int(os.environ["CG"])] if "CG" in os.environ else [1, 2]):
    s = ctypes.c_int(1)
    call("reid_knn_tc_plan", N, N, cg, ctypes.byref(s))
    s = s.value
    keep = k + 34
    tau = torch.empty(N, dtype=torch.int32, device="cuda")
    cand = torch.empty(N * 2 * s * knn_tc.TC_CAP, dtype=torch.int64, device="cuda")
    cnt = torch.zeros(N * 2 * s, dtype=torch.int32, device="cuda")
    ts = []
    for r in range(reps + 2):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        call("reid_knn_candidates_tc", ptr(xh), N, D, knn_tc.SCALE_LOG2, 0, N, keep, s, cg, ptr(cand), ptr(cnt), ptr(tau), stream_ptr())
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = min(ts[2:])
    print("cta_group %d splits %d keep %d: %.3f ms  %.1f TFLOP/s" % (cg, s, keep, t, 2.0 * N * N * D / t / 1e9), flush=True)
