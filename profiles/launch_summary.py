#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: launches, total ms and share per kernel.
Usage: launch_summary.py <launches.csv>"""
import collections, csv, re, sys
rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
scale = {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "s": 1e3, "second": 1e3, "nsecond": 1e-6}
agg = collections.OrderedDict()
for r in rows[1:]:
    name = re.sub(r"<.*", "<...>", re.sub(r"\(.*", "", r[ik]))
    ms = float(r[iv].replace(",", "")) * scale[r[iu]]
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += ms
tot = sum(v[1] for v in agg.values())
print(f"{'kernel':58s} {'launches':>8s} {'total ms':>12s} {'share':>7s}")
for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:58s} {n:8d} {ms:12.3f} {100 * ms / tot:6.2f}%")
print(f"{'TOTAL':58s} {sum(v[0] for v in agg.values()):8d} {tot:12.3f}")
