"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name (shares of the step).
usage: python scripts/launch_shares.py gpurun_out/launches_x.csv [out.md]"""
import csv
import re
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = next(r for r in rows if "Kernel Name" in r)
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = {}
for r in rows:
    if r is hdr or r[ki] == "Kernel Name":
        continue
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0}.get(r[ui], 1e-6)
    name = re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("mip360::", "")
    a = agg.setdefault(name, [0.0, 0])
    a[0] += v
    a[1] += 1
tot = sum(a[0] for a in agg.values())
lines = [f"# launch list shares: {sys.argv[1]} ({sum(a[1] for a in agg.values())} launches, {tot:.2f} ms serialised, cold cache)", "",
         "| kernel | launches | total ms | share |", "|---|---|---|---|"]
for name, (ms, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    lines.append(f"| `{name}` | {n} | {ms:.3f} | {100 * ms / tot:.1f} % |")
text = "\n".join(lines)
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(text + "\n")
print(text)
