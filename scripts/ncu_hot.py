"""Hot SASS lines of an ncu report (needs -lineinfo + --import-source on). Usage: ncu_hot.py rep.ncu-rep [top]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
kn = sys.argv[3] if len(sys.argv) > 3 else None
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"] + (["-k", "regex:" + kn] if kn else []), capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
ci, si, ii = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
data = []
for n, r in enumerate(rows[2:]):
    try:
        data.append((int(r[si] or 0), int(r[ii] or 0), n, r[ci].strip(), r))
    except Exception:
        pass
tot = sum(d[0] for d in data); toti = sum(d[1] for d in data)
print("total samples", tot, "warp inst", toti)
for s, i, n, src, r in sorted(data, key=lambda d: -d[0])[:top]:
    st = sorted(((int(r[c] or 0), h) for c, h in stall_cols), reverse=True)[:2]
    print("%5d %6.2f%% samp %6.2f%% inst  %-70s %s" % (n, 100 * s / tot, 100 * i / max(toti, 1), src[:70], " ".join("%s=%d" % (h[6:], v) for v, h in st if v)))
groups = {"try_wait": "SYNCS.PHASECHK", "LDTM": "LDTM", "REDUX": "REDUX", "STG": "STG", "FSETP": "FSETP", "UTCHMMA": "UTCHMMA", "UTMALDG": "UTMALDG"}
for k, pat in groups.items():
    print(k, "samples %.2f%%" % (100 * sum(d[0] for d in data if pat in d[3]) / tot), "inst", sum(d[1] for d in data if pat in d[3]))
for s, i, n, src, r in data:
    if "SYNCS.PHASECHK" in src:
        print("  wait", src[:80], "samples %.2f%%" % (100 * s / tot), "inst", i)
