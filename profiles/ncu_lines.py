#!/usr/bin/env python
"""Summarise `ncu -i X.ncu-rep --page source --print-source cuda,sass --csv` by CUDA source line:
share of warp-stall samples and executed warp instructions per line (top N)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur = None; hdr = None; data = []
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < 10: continue
    if r[2] != "-": continue          # keep the per-source-line aggregate rows (Address == '-')
    try:
        s = int(r[6]); ie = int(r[7])
    except ValueError:
        continue
    data.append((s, ie, cur, r[0], r[1].strip()[:100]))
tot = sum(d[0] for d in data); toti = sum(d[1] for d in data)
print(f"total samples {tot}  total warp-instructions {toti}")
for s, ie, f, ln, src in sorted(data, reverse=True)[:top]:
    print(f"{100*s/max(tot,1):5.1f}% smp  {100*ie/max(toti,1):5.1f}% inst  {f}:{ln}: {src}")
