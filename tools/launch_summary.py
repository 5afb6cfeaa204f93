"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name (us, share)."""
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
        unit = r.get("Metric Unit", "ns")
        us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
        rows.append((int(r["ID"]), re.sub(r"\(.*", "", r["Kernel Name"]), us))
first = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = [r for r in rows if r[0] >= first]
agg = OrderedDict()
for _, name, us in rows:
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += us
tot = sum(a[1] for a in agg.values())
print(f"# {len(rows)} launches, {tot:.1f} us of kernel time (cold-cache, serialised: compare SHARES)")
for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{100 * us / tot:7.3f}%  n={n:5d}  {us:12.1f} us  avg {us / n:9.2f} us  {name}")
