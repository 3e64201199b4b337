"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel name the
number of launches, total device time and share of the captured step. Usage:
    python tools/summarize_launches.py gpurun_out/launches_r01.csv > profiles/r01_launches_summary.csv"""
import csv
import re
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r["Metric Name"] != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    ns = v * {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1.0)
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"<.*", "", name)
    rows.append((name, ns))
agg = defaultdict(lambda: [0, 0.0])
for n, ns in rows:
    agg[n][0] += 1
    agg[n][1] += ns
total = sum(v[1] for v in agg.values())
print(f"# {len(rows)} launches, {total/1e6:.3f} ms summed device time (cold-cache, serialised under ncu: compare SHARES)")
print("kernel,launches,total_us,share")
for n, (c, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{n},{c},{ns/1e3:.1f},{ns/total:.4f}")
