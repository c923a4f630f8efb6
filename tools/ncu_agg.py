"""Aggregates an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel name.
usage: python tools/ncu_agg.py <launches.csv> [top]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1], errors="ignore")))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
for i, r in enumerate(rows):
    if "Kernel Name" in r:
        hdr, start = r, i
        break
ki, mi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
total = 0.0
for r in rows[start + 1:]:
    if len(r) <= mi or not r[mi].replace(",", "").replace(".", "").isdigit():
        continue
    t = float(r[mi].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0}.get(r[ui], 1e-6)
    a = agg.setdefault(r[ki][:90], [0, 0.0])
    a[0] += 1
    a[1] += t
    total += t
print(f"{sum(a[0] for a in agg.values())} launches, {total:.3f} ms serialised")
for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:top]:
    print(f"{n:5d} {t:9.3f} ms {100 * t / total:5.1f} %  {k}")
