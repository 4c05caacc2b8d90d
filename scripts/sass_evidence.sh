#!/usr/bin/env bash
# Blackwell-specific SASS mnemonics per kernel of the shipped library -> profiles/r02_sass_evidence.txt
# (instruction counts per kernel; kernels without any of them are left out).
so="$(dirname "$0")/../reid-gan_b200/libreid_b200.so"
cuobjdump -sass "$so" | python3 -c '
import collections, re, subprocess, sys
cur, cnt = None, collections.OrderedDict()
pat = re.compile(r"\b(UTC[A-Z]*MMA(?:\.[0-9A-Z]+)*|UTCBAR(?:\.[0-9A-Z]+)*|UTMALDG(?:\.[0-9A-Z]+)*|LDTM(?:\.[0-9a-zA-Z]+)*|DMMA(?:\.[0-9a-zA-Z]+)*|HMMA\.[0-9]+\.F32\.TF32)\b")
for line in sys.stdin:
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    m = pat.search(line)
    if m and cur:
        cnt.setdefault(cur, collections.Counter())[m.group(1)] += 1
names = subprocess.run(["c++filt"] + list(cnt), capture_output=True, text=True).stdout.split("\n")
for mangled, name in zip(cnt, names):
    name = re.sub(r"\(.*", "", name)
    print("%-60s %s" % (name, "  ".join("%s x%d" % kv for kv in sorted(cnt[mangled].items()))))
'
