#!/usr/bin/env python
"""tools/ncu_stalls.py <report.ncu-rep> — warp stall samples of the kernel by reason (source page totals)."""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None; tot = {}
for r in rows:
    if r and r[0] in ("Address", "Line No") or (len(r) > 3 and "Source" in r[:3] and "stall_wait" in r): hdr = r; continue
    if hdr is None or len(r) < len(hdr): continue
    for h, v in zip(hdr, r):
        if h.startswith("stall_") and "Not Issued" not in h and v.isdigit(): tot[h] = tot.get(h, 0) + int(v)
s = sum(tot.values())
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]): print(f"{k:28s} {v:9d} {100*v/max(s,1):5.1f}%")
