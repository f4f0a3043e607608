#!/usr/bin/env python
"""tools/ncu_lines.py <report.ncu-rep> [top] — per source line: warp instructions executed and stall samples."""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur = None; hdr = None; acc = []
for r in rows:
    if len(r) == 2 and r[0] in ("File Path", "File Name"): cur = r[1].split("/")[-1]; continue
    if len(r) == 2: continue
    if r and r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < len(hdr) or not r[0].isdigit(): continue
    d = dict(zip(hdr, r))
    num = lambda x: int(x) if x.lstrip("-").isdigit() else 0
    acc.append((num(d["Instructions Executed"]), num(d["# Samples"]), cur, int(r[0]), r[1].strip()[:110]))
ti = sum(a[0] for a in acc); ts = sum(a[1] for a in acc)
print(f"total warp instructions {ti:,}  samples {ts:,}")
for key, name in ((0, "by instructions"), (1, "by stall samples")):
    print(f"--- top {top} lines {name} ---")
    for a in sorted(acc, key=lambda a: -a[key])[:top]:
        print(f"{100*a[0]/ti:5.1f}% inst {100*a[1]/max(ts,1):5.1f}% smpl  {a[2]}:{a[3]:<4d} {a[4]}")
