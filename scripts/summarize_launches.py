"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (python scripts/summarize_launches.py FILE)."""
import collections
import csv
import re
import sys

lines = [l for l in open(sys.argv[1]) if l.startswith('"')]
agg = collections.defaultdict(lambda: [0, 0.0])
tot = 0.0
for row in csv.DictReader(lines):
    name = row["Kernel Name"]
    m = re.search(r"(\w+)(<[^>]*>)?\(", name)
    k = (m.group(1) + (m.group(2) or "")) if m else name
    k += " grid=" + row["Grid Size"] + " block=" + row["Block Size"]
    t = float(row["Metric Value"]) / 1e3
    agg[k][0] += 1
    agg[k][1] += t
    tot += t
print(f"total {tot / 1e3:.2f} ms over {sum(v[0] for v in agg.values())} launches (ncu per-launch times: cold-cache, serialised)")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{v[1]:10.1f} us {v[0]:5d}x {v[1] / tot * 100:5.1f}%  {k}")
