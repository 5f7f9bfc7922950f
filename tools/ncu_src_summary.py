#!/usr/bin/env python
"""Summarise `ncu --page source --print-source cuda,sass --csv` by CUDA source line:
warp instructions executed and stall samples, top N lines per file."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur = None
agg = []
hdr = None
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No":
        hdr = r; continue
    if hdr is None or not r or r[0] == "" or r[0] in ("Function Name",):
        continue
    try:
        ln = int(r[0])
    except ValueError:
        continue
    d = dict(zip(hdr[2:], r[2:]))
    def num(k):
        try:
            return int(d.get(k, "0") or 0)
        except ValueError:
            return 0
    inst = num("Instructions Executed")
    samp = num("# Samples")
    agg.append((cur, ln, r[1].strip()[:90], inst, samp))
ti = sum(a[3] for a in agg) or 1
ts = sum(a[4] for a in agg) or 1
print(f"total warp instructions {ti}, samples {ts}")
print("--- by instructions")
for a in sorted(agg, key=lambda a: -a[3])[:top]:
    print(f"{100*a[3]/ti:5.1f}% inst {100*a[4]/ts:5.1f}% samp  {a[0]}:{a[1]}  {a[2]}")
print("--- by stall samples")
for a in sorted(agg, key=lambda a: -a[4])[:top]:
    print(f"{100*a[4]/ts:5.1f}% samp {100*a[3]/ti:5.1f}% inst  {a[0]}:{a[1]}  {a[2]}")
