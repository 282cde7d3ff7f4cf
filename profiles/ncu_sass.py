#!/usr/bin/env python
"""SASS rows (with sample counts and the dominant stall reasons) of given CUDA source lines of an ncu capture.
usage: ncu_sass.py <rep> <file-substring> <line_lo> <line_hi>"""
import csv, subprocess, sys
rep, fsub, lo, hi = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file, hdr, on = None, None, False
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or r[0] == "Function Name": continue
    if r[0].isdigit() and r[2] == "-":
        on = (fsub in (cur_file or "")) and lo <= int(r[0]) <= hi
        if on: print(f"--- {r[0]}: {r[1].strip()[:120]}")
        continue
    if on and len(r) > 3 and r[2] != "-":
        d = dict(zip(hdr, r))
        try: smp = int(float(d.get("# Samples", "0") or 0))
        except ValueError: smp = 0
        stalls = []
        for k, v in d.items():
            if k.startswith("stall_") and "(Not Issued)" not in k:
                try:
                    if float(v) > 0: stalls.append((float(v), k[6:]))
                except ValueError: pass
        stalls.sort(reverse=True)
        print(f"   {smp:6d}  inst {d.get('Instructions Executed','')[:9]:>9}  {r[3].strip()[:60]:60s} " + " ".join(f"{n}:{int(v)}" for v, n in stalls[:4]))
