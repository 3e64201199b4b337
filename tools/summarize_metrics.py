"""Summarises an `ncu --metrics ... --csv` capture (long format) per launch and per kernel:
duration, DRAM bytes, achieved DRAM GB/s, tensor-pipe and DRAM utilisation.
    python tools/summarize_metrics.py gpurun_out/fwd_metrics_r01.csv [--per-launch]"""
import csv
import re
import sys
from collections import OrderedDict, defaultdict

UNIT = {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "%": 1.0,
        "register/thread": 1.0}
launches = OrderedDict()
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    k = int(r["ID"])
    d = launches.setdefault(k, {"name": re.sub(r"<.*|\(.*", "", re.sub(r"^void ", "", r["Kernel Name"])), "grid": r["Grid Size"]})
    try:
        d[r["Metric Name"]] = float(r["Metric Value"].replace(",", "")) * UNIT.get(r["Metric Unit"], 1.0)
    except ValueError:
        pass
per_launch = "--per-launch" in sys.argv
if per_launch:
    print("id,kernel,grid,dur_us,dram_read_MB,dram_write_MB,dram_GBps,dram_pct,tensor_pct,sm_pct,regs")
agg = defaultdict(lambda: [0, 0.0, 0.0, 0.0, 0.0, 0.0])
for k, d in launches.items():
    dur = d.get("gpu__time_duration.sum", 0.0)
    rd, wr = d.get("dram__bytes_read.sum", 0.0), d.get("dram__bytes_write.sum", 0.0)
    if per_launch:
        print(f'{k},{d["name"]},"{d["grid"]}",{dur*1e6:.1f},{rd/1e6:.2f},{wr/1e6:.2f},{(rd+wr)/dur/1e9 if dur else 0:.0f},'
              f'{d.get("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 0):.1f},'
              f'{d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0):.1f},'
              f'{d.get("sm__throughput.avg.pct_of_peak_sustained_elapsed", 0):.1f},{d.get("launch__registers_per_thread", 0):.0f}')
    a = agg[d["name"]]
    a[0] += 1; a[1] += dur; a[2] += rd; a[3] += wr
    a[4] += d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0) * dur
    a[5] += d.get("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 0) * dur
if not per_launch:
    tot = sum(a[1] for a in agg.values())
    print(f"# {len(launches)} launches, {tot*1e3:.3f} ms summed (serialised under ncu, cold caches: compare SHARES)")
    print("kernel,launches,total_us,share,dram_read_MB,dram_write_MB,dram_GBps,avg_dram_pct,avg_tensor_pct")
    for n, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{n},{a[0]},{a[1]*1e6:.1f},{a[1]/tot:.4f},{a[2]/1e6:.1f},{a[3]/1e6:.1f},{(a[2]+a[3])/a[1]/1e9 if a[1] else 0:.0f},"
              f"{a[5]/a[1] if a[1] else 0:.1f},{a[4]/a[1] if a[1] else 0:.1f}")
