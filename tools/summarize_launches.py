"""Summarise an ncu launch-list CSV (gpu__time_duration.sum) per kernel name: count, total ms, share."""
import csv
import sys
from collections import defaultdict

path = sys.argv[1]
rows = []
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
r = csv.DictReader(lines)
agg = defaultdict(lambda: [0, 0.0])
total = 0.0
for row in r:
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = row["Kernel Name"].split("(")[0]
    v = float(row["Metric Value"].replace(",", ""))
    unit = row.get("Metric Unit", "ns")
    ms = v / 1e6 if unit in ("ns", "nsecond") else (v / 1e3 if unit in ("us", "usecond") else v)
    agg[name][0] += 1
    agg[name][1] += ms
    total += ms
print(f"total {total:.2f} ms over {sum(a[0] for a in agg.values())} launches")
for name, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{ms:9.3f} ms  {100 * ms / total:5.1f}%  x{n:4d}  {name}")
