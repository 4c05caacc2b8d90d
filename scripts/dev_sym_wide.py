"""A/B of the symmetric candidate kernel: single tiles (reid_knn_candidates_sym) against 256 x 512 strips
(reid_knn_candidates_sym_wide): identical candidate sets, CUDA-event time of the main-pass launch alone."""
import os
import sys
sys.path.insert(0, ".")
import torch
import reid_gan_b200 as rg
from reid_gan_b200 import _lib, knn_tc as kt
DBG = int(os.environ.get("REID_TC_DEBUG", "0"))     # developer switches need the -DREID_DEV build (REID_DEV=1 csrc/build.sh)
if DBG:
    _lib.LIB_PATH = _lib.LIB_PATH.replace("libreid_b200.so", "libreid_b200_dev.so")
from reid_gan_b200._lib import call, ptr, stream_ptr


def run(N, D, n_ids, reps=10):
    x, _ = rg.synth(N, D, n_ids, 0.8, 0)
    x = x.cuda()
    dev = x.device
    sp = stream_ptr()
    xh = torch.empty((N, D), dtype=torch.float16, device=dev)
    msq = torch.zeros(2, device=dev)
    call("reid_features_to_half", ptr(x), N, D, kt.SCALE_LOG2, ptr(xh), ptr(msq), sp)
    kt.SYM_WIDE = False
    cand, cnt, tau_ord, cap, info = kt._candidates_sym(xh, N, D, sp, dev)
    tau = torch.empty(N, dtype=torch.float32, device=dev)
    # thresholds again (the helper does not return the float ones): same calls as _candidates_sym
    m = kt.sample_size(N)
    xs = torch.empty((m, D), dtype=torch.float16, device=dev)
    call("reid_features_sample", ptr(xh), N, D, m, kt._sample_stride(N, m), ptr(xs), sp)
    pre = torch.empty(N * 2 * kt.TC_CAP, dtype=torch.int64, device=dev)
    pre_cnt = torch.zeros(N * 2, dtype=torch.int32, device=dev)
    pre_tau = torch.empty(N, dtype=torch.int32, device=dev)
    call("reid_knn_candidates_tc_ab", ptr(xh), N, ptr(xs), m, D, kt.SCALE_LOG2, 0, N, -kt.sym_rank(30), 1, 2, ptr(pre), ptr(pre_cnt),
         ptr(pre_tau), sp)
    t_o = torch.empty(N, dtype=torch.int32, device=dev)
    call("reid_knn_sample_tau", ptr(pre), ptr(pre_cnt), ptr(pre_tau), 2, N, kt.sym_rank(30), ptr(tau), ptr(t_o), sp)
    tiles = kt._tile_order((N + 255) // 256, dev)
    res = {}
    for wide in (False, True):
        kt.SYM_WIDE = wide
        c = torch.zeros(N * cap, dtype=torch.int64, device=dev)
        n = torch.empty(N, dtype=torch.int32, device=dev)
        for _ in range(2):
            kt.candidates_sym_launch(xh, N, D, tau, tiles, cap, c, n, 1, sp)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            kt.candidates_sym_launch(xh, N, D, tau, tiles, cap, c, n, 1, sp)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        c2 = c.view(N, cap).clone()
        nn = n.clamp(max=cap).long()
        mask = torch.arange(cap, device=dev)[None, :] >= nn[:, None]
        c2[mask] = torch.iinfo(torch.int64).max
        res[wide] = (n.clone(), torch.sort(c2, dim=1).values, ms)
        n_t = tiles.shape[0]
        print("N=%d wide=%s: %.4f ms/launch, %.0f TFLOP/s executed, mean list %.1f, max %d" % (
            N, wide, ms, 2.0 * 256 * 256 * D * n_t / (ms * 1e-3) / 1e12, float(n.float().mean()), int(n.max())), flush=True)
    same_n = bool(torch.equal(res[False][0], res[True][0]))
    same_c = bool(torch.equal(res[False][1], res[True][1]))
    print("N=%d counts equal %s, candidate sets equal %s, speed-up %.3f" % (N, same_n, same_c, res[False][2] / res[True][2]), flush=True)
    assert DBG or (same_n and same_c)


if __name__ == "__main__":
    if not DBG:
        run(8192 + 77, 256, 300, reps=3)
        run(12936, 2048, 751, reps=5)
    run(32621, 2048, 1041, reps=10)
