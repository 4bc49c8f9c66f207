#!/usr/bin/env python
"""Summarise an `ncu --page source --csv` dump: stall-reason totals and the hottest SASS lines."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
h = rows[1]
d = [r for r in rows[2:] if len(r) == len(h) and r[0] != "Address"]
ci = {n: i for i, n in enumerate(h)}
samp = ci["# Samples"]
stalls = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
tot = {s: 0 for s in stalls}
total = 0
inst = 0
for r in d:
    try:
        total += int(r[samp]); inst += int(r[ci["Instructions Executed"]])
    except ValueError:
        continue
    for s in stalls:
        tot[s] += int(r[ci[s]] or 0)
print("samples", total, "warp-instructions", inst)
for s, v in sorted(tot.items(), key=lambda kv: -kv[1])[:10]:
    print(f"  {s:28s} {v:8d} {100*v/max(total,1):5.1f}%")
print("hottest lines:")
for r in sorted(d, key=lambda r: -int(r[samp] or 0))[:topn]:
    top = sorted(((int(r[ci[s]] or 0), s) for s in stalls), reverse=True)[:2]
    print(f"  {int(r[samp]):6d} {100*int(r[samp])/total:4.1f}% exec={r[ci['Instructions Executed']]:>9s} {r[ci['Source']].strip()[:70]:70s} {top[0][1][6:]}:{top[0][0]} {top[1][1][6:]}:{top[1][0]}")
