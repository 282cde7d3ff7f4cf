#!/usr/bin/env python
"""All stall reasons of the SASS attributed to given source lines: ncu_line_stalls.py <rep> <file> <first> <last>"""
import csv, subprocess, sys, collections, os
rep, f, lo, hi = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
cur = None; hdr = None; agg = collections.Counter(); n = 0; inst = 0
for r in csv.reader(out.splitlines()):
    if not r: continue
    if r[0] == "File Path": cur = os.path.basename(r[1]); continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or not r[0].isdigit() or r[2] != "-": continue
    if cur != f or not (lo <= int(r[0]) <= hi): continue
    d = dict(zip(hdr, r))
    for k, v in d.items():
        if k.startswith("stall_") and "Not Issued" not in k:
            try: agg[k] += float(v or 0)
            except ValueError: pass
    n += float(d.get("# Samples") or 0); inst += float(d.get("Instructions Executed") or 0)
print(f"{f}:{lo}-{hi}: {n:.0f} samples, {inst:.0f} warp instructions")
print("  " + "  ".join(f"{k[6:]} {100 * v / max(n, 1):.1f}%" for k, v in agg.most_common(9)))
