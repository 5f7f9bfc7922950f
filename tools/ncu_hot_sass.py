#!/usr/bin/env python
"""Hot path in address order from `ncu --page source --print-source sass --csv`: executed count per
unit (argv[2] = number of units, e.g. warp-tiles), stall samples, instruction."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
units = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
thr = float(sys.argv[3]) if len(sys.argv) > 3 else 0.05
hdr = None
tot = 0
for r in rows:
    if r and r[0] == "Address":
        hdr = r
        continue
    if hdr and r and r[0].startswith("0x"):
        d = dict(zip(hdr, r))
        n = int(d["Instructions Executed"] or 0)
        tot += n
        if n / units >= thr:
            print(f"{n / units:7.2f} {int(d['# Samples'] or 0):6d}  {d['Source'].strip()}")
print("total per unit", tot / units)
