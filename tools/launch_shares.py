#!/usr/bin/env python
"""Per-kernel share of GPU time from an ncu launch list:
   ncu --metrics gpu__time_duration.sum --clock-control none -c N --csv --log-file list.csv <command>
   python tools/launch_shares.py list.csv"""
import csv, sys
from collections import defaultdict
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = next(r for r in rows if "Kernel Name" in r)
ik, iv, im = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
iu = hdr.index("Metric Unit")
tot = defaultdict(float); cnt = defaultdict(int)
for r in rows:
    if r is hdr or r[im] != "gpu__time_duration.sum":
        continue
    v = float(r[iv].replace(",", "")); u = r[iu]
    ms = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, 1e-6)
    tot[r[ik]] += ms; cnt[r[ik]] += 1
s = sum(tot.values())
for k in sorted(tot, key=lambda k: -tot[k]):
    print("%-70s launches=%4d total_ms=%9.3f share=%.3f" % (k[:70], cnt[k], tot[k], tot[k] / s))
