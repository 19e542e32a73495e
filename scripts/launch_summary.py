#!/usr/bin/env python
"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list.
Usage: python scripts/launch_summary.py <launches.csv> <steps in the capture> [out.txt]"""
import collections
import csv
import sys

src, steps = sys.argv[1], int(sys.argv[2])
lines = [l for l in open(src) if l.startswith('"')]
rows = list(csv.reader(lines))
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[1:]:
    agg.setdefault(r[ki], []).append(float(r[vi].replace(",", "")))
tot = sum(sum(v) for v in agg.values()) / steps / 1000
out = [f"ncu --metrics gpu__time_duration.sum --clock-control none, NVTX range spmf_timed ({steps} steps); "
       "times are serialised cold-cache launches: shares, not absolutes"]
for n, v in agg.items():
    out.append(f"{n[:96]:96s} n/step={len(v) / steps:5.1f} avg={sum(v) / len(v) / 1000:9.1f} us "
               f"per-step={sum(v) / steps / 1000:8.1f} us share={sum(v) / steps / 1000 / tot:6.1%}")
out.append(f"total per step {tot:.1f} us (serialised)")
text = "\n".join(out) + "\n"
if len(sys.argv) > 3:
    open(sys.argv[3], "w").write(text)
print(text)
