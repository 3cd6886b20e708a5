#!/usr/bin/env python
"""Top stall sites of an `ncu --page source --csv` export (SASS view): tools/ncu_src_top.py file.csv [n]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = 0
recs = []
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    try:
        s = int(r[ix["# Samples"]] or 0)
    except ValueError:
        continue
    tot += s
    top = sorted(((int(r[ix[h]] or 0), h) for h in stalls), reverse=True)[:2]
    recs.append((s, r[ix["Address"]], r[ix["Source"]][:90], top))
print("kernel:", rows[0][1], "total samples", tot)
for s, a, src, top in sorted(recs, reverse=True)[:n]:
    print(f"{100*s/max(tot,1):5.1f}% {a[-5:]} {src:90s} {top}")
