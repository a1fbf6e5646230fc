#!/usr/bin/env python
"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel: total ms, share, launches, mean us.
    python tools/launch_summary.py gpurun_out/r1_launches.csv > profiles/r1_launches_summary.txt"""
import collections, csv, re, sys
rows = list(csv.reader(open(sys.argv[1], errors="replace")))
start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
hdr = rows[start]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
n = 0
for r in rows[start + 1:]:
    if len(r) <= vi:
        continue
    name = re.sub(r"\(.*$", "", r[ki]).replace("void ", "").replace("<unnamed>::", "").replace("(anonymous namespace)::", "").strip()
    t = float(r[vi].replace(",", ""))
    t_us = t / 1e3 if r[ui].startswith("ns") or r[ui] == "nsecond" else (t if r[ui].startswith("us") else t * 1e3)
    a = agg.setdefault(name, [0.0, 0])
    a[0] += t_us; a[1] += 1; n += 1
tot = sum(v[0] for v in agg.values())
print(f"# {n} launches, sum {tot / 1e3:.2f} ms")
print("#        ms   share launches    avg us  kernel")
for name, (t, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"{t / 1e3:10.3f} {100 * t / tot:6.2f}% {c:8d} {t / c:9.1f}  {name[:110]}")
