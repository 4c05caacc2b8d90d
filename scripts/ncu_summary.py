"""Summarise an .ncu-rep (read here, no GPU): per kernel the metrics the roofline discussion needs.
Usage: python scripts/ncu_summary.py gpurun_out/prof.ncu-rep [kernel-regex]"""
import csv
import re
import subprocess
import sys

rep = sys.argv[1]
pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h, units = rows[0], rows[1]
WANT = [
    ("time_us", "gpu__time_duration.sum"),
    ("dramR_MB", "dram__bytes_read.sum"), ("dramW_MB", "dram__bytes_write.sum"),
    ("dram%", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    ("lts%", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("l1tex%", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("lsu_wave%", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
    ("sm%", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("issue%", "sm__inst_issued.avg.pct_of_peak_sustained_active"),
    ("tensor%", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"),
    ("fp64%", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed"),
    ("fp64rt%", "TPC.TriageCompute.sm__pipe_fp64_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed"),
    ("xu%", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
    ("alu%", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
    ("occ%", "sm__warps_active.avg.pct_of_peak_sustained_active"),
    ("L2hit%", "lts__t_sector_hit_rate.pct"),
    ("regs", "launch__registers_per_thread"), ("grid", "launch__grid_size"), ("block", "launch__block_size"),
    ("smem_dyn", "launch__shared_mem_per_block_dynamic"), ("smem_st", "launch__shared_mem_per_block_static"),
]
STALL = [c for c in h if "smsp__average_warp" in c and "issue_stalled" in c and c.endswith("_per_warp_active.pct")] or \
        [c for c in h if "smsp__average_warps_issue_stalled" in c and c.endswith(".ratio")]
ki = h.index("Kernel Name")
for r in rows[2:]:
    d = dict(zip(h, r))
    name = d["Kernel Name"].split("(")[0]
    if pat and not pat.search(name):
        continue
    print("==", name)
    out = []
    for label, col in WANT:
        if col in d and d[col] != "":
            u = units[h.index(col)]
            out.append("%s=%s%s" % (label, d[col], "" if label.endswith("%") or u in ("", "%") else u))
    print("  ", "  ".join(out))
    st = []
    for c in STALL:
        try:
            v = float(d[c])
        except ValueError:
            continue
        st.append((v, c.replace("smsp__average_warps_issue_stalled_", "").replace("smsp__average_warp_latency_issue_stalled_", "").split("_per_")[0].replace(".ratio", "")))
    st.sort(reverse=True)
    print("   stalls:", "  ".join("%s=%.2f" % (n, v) for v, n in st[:7]))
