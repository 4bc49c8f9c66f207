#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv --log-file x.csv` launch list per kernel:
    python tools/ncu_launches.py gpurun_out/x.csv > profiles/x_summary.csv"""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot, cnt = collections.OrderedDict(), collections.Counter()
for r in rows[1:]:
    if r[kn] == "Kernel Name":
        continue
    name = r[kn].split("(")[0].replace("void ", "")[:90]
    v = float(r[mv].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[mu], 1e-3)
    tot[name] = tot.get(name, 0.0) + v
    cnt[name] += 1
w = csv.writer(sys.stdout)
w.writerow(["kernel", "launches", "avg_us", "total_us"])
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    w.writerow([k, cnt[k], round(v / cnt[k], 2), round(v, 1)])
