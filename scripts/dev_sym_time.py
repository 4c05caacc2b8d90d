"""Time the symmetric candidate kernel alone (developer switches via REID_TC_DEBUG)."""
import sys, os
sys.path.insert(0, ".")
import torch
import reid_gan_b200 as rg
from reid_gan_b200 import knn_tc, _lib
from reid_gan_b200._lib import call, ptr, stream_ptr
N, D = 32621, 2048
x, _ = rg.synth(N, D, 1041, 0.8, 0)
x = x.cuda()
xh = torch.empty((N, D), dtype=torch.float16, device="cuda")
msq = torch.zeros(1, device="cuda")
call("reid_features_to_half", ptr(x), N, D, 4, ptr(xh), ptr(msq), stream_ptr())
for sb in [int(v) for v in os.environ.get("SB", "8").split(",")]:
    knn_tc._tile_cache.clear()
    tiles = knn_tc._tile_order((N + 255) // 256, x.device, sb)
    for _ in range(2):
        out = knn_tc._candidates_sym(xh, N, D, stream_ptr(), x.device)
    _lib.profiler.start()
    for _ in range(5):
        out = knn_tc._candidates_sym(xh, N, D, stream_ptr(), x.device)
    torch.cuda.synchronize()
    _lib.profiler.stop()
    print("dbg", os.environ.get("REID_TC_DEBUG", "0"), "sb", sb, {k: round(v[1] / v[0], 4) for k, v in _lib.profiler.summary().items()}, "cand mean", float(out[1].float().mean()))
