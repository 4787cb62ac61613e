"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel launches, total ns and share.
    python tools/launch_summary.py gpurun_out/launches.csv [steps_in_capture]"""
import collections
import csv
import re
import sys

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]
ki, vi, mi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
ui = hdr.index("Metric Unit")
agg = collections.OrderedDict()
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1.0, "us": 1e3, "ms": 1e6}
for r in rows[1:]:
    name = re.sub(r"\(.*", "", r[ki])
    name = re.sub(r"^void ", "", name).replace("<unnamed>::", "")
    a = agg.setdefault(name, [0, 0.0, 0.0])
    val = float(r[vi].replace(",", "")) * SCALE.get(r[ui], 1.0)
    if r[mi] == "gpu__time_duration.sum":
        a[0] += 1
        a[1] += val
    elif r[mi].startswith("dram__bytes"):
        a[2] += val
tot = sum(v[1] for v in agg.values())
nl = sum(v[0] for v in agg.values())
print(f"# {nl} launches in the capture ({steps} steps), {tot / 1e6 / steps:.3f} ms of kernel time per step (cold-cache, serialised under ncu)")
print(f"{'kernel':70s} {'launches/step':>13s} {'us/step':>10s} {'share':>7s} {'DRAM MB/launch':>15s}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:70]:70s} {v[0] / steps:13.1f} {v[1] / 1e3 / steps:10.1f} {v[1] / tot * 100:6.1f}% {v[2] / max(v[0], 1) / 1e6:15.2f}")
