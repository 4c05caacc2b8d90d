"""What bounds the symmetric candidate kernel: SM clock, board power and throttle reasons sampled (NVML, 20 ms) while the
kernel alone runs back to back for a few seconds -- in the release flavour and, with REID_TC_DEBUG (developer build),
with parts switched off (1 = no TMEM reads in the epilogue, 4 = no TMA: MMAs on whatever the ring holds, 5 = both).
Prints ms/launch, executed TFLOP/s, and the clock / power statistics of the loop."""
import os
import sys
import threading
import time
sys.path.insert(0, ".")
import torch
import pynvml
import reid_gan_b200 as rg
from reid_gan_b200 import _lib, knn_tc as kt
DBG = int(os.environ.get("REID_TC_DEBUG", "0"))
if DBG:
    _lib.LIB_PATH = _lib.LIB_PATH.replace("libreid_b200.so", "libreid_b200_dev.so")
from reid_gan_b200._lib import call, ptr, stream_ptr

N, D = 32621, 2048
SECONDS = float(sys.argv[1]) if len(sys.argv) > 1 else 3.0
x, _ = rg.synth(N, D, 1041, 0.8, 0)
x = x.cuda()
dev = x.device
sp = stream_ptr()
xh = torch.empty((N, D), dtype=torch.float16, device=dev)
msq = torch.zeros(2, device=dev)
call("reid_features_to_half", ptr(x), N, D, kt.SCALE_LOG2, ptr(xh), ptr(msq), sp)
cand, cnt, tau_ord, cap, info = kt._candidates_sym(xh, N, D, sp, dev)
m = kt.sample_size(N)
xs = torch.empty((m, D), dtype=torch.float16, device=dev)
call("reid_features_sample", ptr(xh), N, D, m, kt._sample_stride(N, m), ptr(xs), sp)
pre = torch.empty(N * 2 * kt.TC_CAP, dtype=torch.int64, device=dev)
pre_cnt = torch.zeros(N * 2, dtype=torch.int32, device=dev)
pre_tau = torch.empty(N, dtype=torch.int32, device=dev)
call("reid_knn_candidates_tc_ab", ptr(xh), N, ptr(xs), m, D, kt.SCALE_LOG2, 0, N, -kt.sym_rank(30), 1, 2, ptr(pre), ptr(pre_cnt),
     ptr(pre_tau), sp)
tau = torch.empty(N, dtype=torch.float32, device=dev)
t_o = torch.empty(N, dtype=torch.int32, device=dev)
call("reid_knn_sample_tau", ptr(pre), ptr(pre_cnt), ptr(pre_tau), 2, N, kt.sym_rank(30), ptr(tau), ptr(t_o), sp)
tiles = kt._tile_order((N + 255) // 256, dev)
kt.SYM_WIDE = os.environ.get("WIDE", "0") == "1"         # 256 x 512 strips instead of single tiles
torch.cuda.synchronize()

pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(torch.cuda.current_device())
samples = []
stop = False


def poll():
    while not stop:
        samples.append((pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetPowerUsage(h) / 1e3,
                        pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)))
        time.sleep(0.02)


for _ in range(3):
    kt.candidates_sym_launch(xh, N, D, tau, tiles, cap, cand, cnt, 1, sp)
torch.cuda.synchronize()
th = threading.Thread(target=poll)
th.start()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.time()
n = 0
e0.record()
while time.time() - t0 < SECONDS:
    for _ in range(50):
        kt.candidates_sym_launch(xh, N, D, tau, tiles, cap, cand, cnt, 1, sp)
    n += 50
    torch.cuda.synchronize()
e1.record()
torch.cuda.synchronize()
stop = True
th.join()
ms = e0.elapsed_time(e1) / n
half = samples[len(samples) // 2:]                       # steady state: second half of the loop
clk = sorted(s[0] for s in half)
pw = sorted(s[1] for s in half)
reasons = 0
for s in half:
    reasons |= s[2]
names = {pynvml.nvmlClocksEventReasonSwPowerCap: "sw_power_cap", pynvml.nvmlClocksEventReasonHwSlowdown: "hw_slowdown",
         pynvml.nvmlClocksEventReasonSwThermalSlowdown: "sw_thermal_slowdown",
         pynvml.nvmlClocksEventReasonHwThermalSlowdown: "hw_thermal_slowdown",
         pynvml.nvmlClocksEventReasonHwPowerBrakeSlowdown: "hw_power_brake"}
print("dbg %d wide %d: %.4f ms/launch over %d launches (%.1f s), %.0f TFLOP/s executed; SM clock min/median/max %d/%d/%d MHz (max %d), "
      "power median %.0f W max %.0f W (limit %.0f W), reasons %s" % (
          DBG, int(kt.SYM_WIDE), ms, n, SECONDS, 2.0 * 256 * 256 * D * tiles.shape[0] / (ms * 1e-3) / 1e12, clk[0], clk[len(clk) // 2], clk[-1],
          pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM), pw[len(pw) // 2], pw[-1],
          pynvml.nvmlDeviceGetPowerManagementLimit(h) / 1e3, [v for k_, v in names.items() if reasons & k_]), flush=True)
