"""Summarises an `ncu --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --csv` launch
list by kernel (python scripts/summarize_launches.py FILE [--json OUT]).  With the DRAM metrics present it also
prints the average DRAM traffic per launch of each kernel (the `roofline.traffic` figure of bench.py)."""
import collections
import csv
import json
import re
import sys

lines = [l for l in open(sys.argv[1]) if l.startswith('"')]
per_launch = collections.OrderedDict()
for row in csv.DictReader(lines):
    d = per_launch.setdefault(row["ID"], {"name": row["Kernel Name"], "grid": row["Grid Size"], "block": row["Block Size"]})
    v = float(row["Metric Value"].replace(",", ""))
    unit = row["Metric Unit"]
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "byte": 1.0, "Kbyte": 1e3,
             "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
    d[row["Metric Name"]] = v * scale
agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
tot = 0.0
for d in per_launch.values():
    m = re.search(r"(\w+)(<[^>]*>)?\(", d["name"])
    k = (m.group(1) + (m.group(2) or "")) if m else d["name"]
    k += " grid=" + d["grid"] + " block=" + d["block"]
    t = d.get("gpu__time_duration.sum", 0.0)
    agg[k][0] += 1
    agg[k][1] += t
    agg[k][2] += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
    tot += t
print(f"total {tot / 1e3:.2f} ms over {sum(v[0] for v in agg.values())} launches (ncu per-launch times: cold-cache, serialised)")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    extra = f"  dram {v[2] / v[0] / 1e6:8.1f} MB/launch" if v[2] else ""
    print(f"{v[1]:10.1f} us {v[0]:5d}x {v[1] / tot * 100:5.1f}%  {k}{extra}")
if "--json" in sys.argv:
    g = [v for k, v in agg.items() if k.startswith("gemm_bf16_tcgen05_kernel")]
    n = sum(v[0] for v in g)
    out = {"kernel": "gemm_bf16_tcgen05_kernel", "launches": n, "dram_bytes_per_launch": sum(v[2] for v in g) / max(n, 1),
           "how": "sum of dram__bytes_read.sum + dram__bytes_write.sum over every GEMM launch of one bs=64 step / launches"}
    json.dump(out, open(sys.argv[sys.argv.index("--json") + 1], "w"), indent=1)
    print(out)
