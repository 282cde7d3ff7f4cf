#!/usr/bin/env python
"""Per-source-line summary of an ncu capture: usage ncu_lines.py <rep> [top]
Runs `ncu -i <rep> --page source --print-source cuda,sass --csv` and prints, per CUDA source line, the share of warp
samples, executed warp-instructions and shared-memory wavefronts."""
import csv, subprocess, sys, collections, os
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 45
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file = None; hdr = None
agg = collections.defaultdict(lambda: [0, 0, 0, 0, ""])
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = os.path.basename(r[1]); continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or not r[0].isdigit(): continue
    if r[2] != "-": continue                      # SASS rows carry an address; CUDA rows have '-'
    d = dict(zip(hdr, r))
    def num(k):
        try: return float(d.get(k, "0") or 0)
        except ValueError: return 0.0
    key = (cur_file, int(r[0]))
    a = agg[key]
    a[0] += num("# Samples"); a[1] += num("Instructions Executed"); a[2] += num("L1 Wavefronts Shared"); a[3] += num("L1 Wavefronts Shared Excessive"); a[4] = r[1].strip()[:110]
ts = sum(a[0] for a in agg.values()) or 1; ti = sum(a[1] for a in agg.values()) or 1; tw = sum(a[2] for a in agg.values()) or 1
print(f"total samples {ts:.0f}  warp-instructions {ti:.0f}  shared wavefronts {tw:.0f} (excessive {sum(a[3] for a in agg.values()):.0f})")
for key, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100*a[0]/ts:5.1f}% smp {100*a[1]/ti:5.1f}% inst {100*a[2]/tw:5.1f}% smemwf (exc {100*a[3]/tw:4.1f}%)  {key[0]}:{key[1]}: {a[4]}")
