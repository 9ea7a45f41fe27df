#!/usr/bin/env python3
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: count, total ms, share.
usage: summarize_launches.py launches.csv [title] > profiles/rNN_xxx.txt"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1], errors="ignore")))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr, data = rows[hi], rows[hi + 1:]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
gi, bi = hdr.index("Grid Size"), hdr.index("Block Size")
agg = collections.OrderedDict()
for r in data:
    if len(r) <= vi:
        continue
    name = r[ki].split("(")[0].replace("void ", "")
    v = float(r[vi].replace(",", ""))
    v = {"ns": v / 1e6, "us": v / 1e3, "ms": v, "s": v * 1e3}.get(r[ui], v)
    a = agg.setdefault(name, [0, 0.0, r[gi], r[bi]])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
print("# " + (sys.argv[2] if len(sys.argv) > 2 else sys.argv[1]))
print("# per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes")
print(f"# {'total ms':>10} {'launches':>8} {'ms/launch':>10} {'share':>6}  kernel  (grid, block of first launch)")
for k, (c, t, g, b) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"  {t:10.3f} {c:8d} {t / c:10.3f} {100 * t / tot:5.1f}%  {k[:80]}  {g} {b}")
print(f"  {tot:10.3f} total")
