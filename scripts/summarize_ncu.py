"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (share of the step),
and optionally the key metrics of one `ncu --set full` report exported with `--page raw --csv`.
Usage: python scripts/summarize_ncu.py launches.csv [raw.csv]"""
import csv
import sys
from collections import OrderedDict


def launches(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = OrderedDict()
    for r in data:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1.0)
        name = r[ki].split("(")[0]
        c, t = agg.get(name, (0, 0.0))
        agg[name] = (c + 1, t + v)
    tot = sum(t for _, t in agg.values())
    print("# %s: %d launches, %.1f us in kernels (cold-cache, serialised: compare SHARES)" % (path, sum(c for c, _ in agg.values()), tot))
    print("%-58s %6s %12s %10s %7s" % ("kernel", "calls", "total_us", "avg_us", "share"))
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-58s %6d %12.1f %10.1f %6.1f%%" % (k[:58], c, t, t / c, 100 * t / tot))


KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "lts__t_sectors_srcunit_tex_op_read.sum",
        "sm__warps_active.avg.per_cycle_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__cycles_elapsed.avg.per_second", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "launch__shared_mem_per_block_dynamic"]


def raw(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    for vals in rows[2:]:
        print("# %s :: %s" % (path, vals[ki][:80]))
        for h, u, v in zip(hdr, units, vals):
            if h in KEYS:
                print("%-70s %-12s %s" % (h, u, v))


if __name__ == "__main__":
    launches(sys.argv[1])
    if len(sys.argv) > 2:
        raw(sys.argv[2])
