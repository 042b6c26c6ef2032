"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count/total/avg (us) and share.
usage: python tools/summarize_launches.py file.csv [skip_first_n_launches]"""
import csv
import re
import sys
from collections import OrderedDict

rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
for r in csv.DictReader(lines):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        us = v / 1000.0 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1000.0)
        name = re.sub(r"\(.*", "", r["Kernel Name"])
        rows.append((name, us, r.get("Grid Size", ""), r.get("Block Size", "")))
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = rows[skip:]
agg = OrderedDict()
for name, us, grid, block in rows:
    key = name
    a = agg.setdefault(key, [0, 0.0, set()])
    a[0] += 1
    a[1] += us
    a[2].add(grid)
total = sum(a[1] for a in agg.values())
print(f"{len(rows)} launches, {total:.1f} us total (cold-cache, serialised: compare SHARES)")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{a[1] / total * 100:6.2f}%  n={a[0]:4d}  avg={a[1] / a[0]:9.2f} us  total={a[1]:10.1f} us  {k}  grids={sorted(a[2])[:4]}")
