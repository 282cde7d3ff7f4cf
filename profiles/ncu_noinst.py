#!/usr/bin/env python
"""Where a given stall reason is sampled, per CUDA source line: ncu_noinst.py <rep> [column=stall_no_inst] [top=30]"""
import csv, subprocess, sys, collections, os
rep = sys.argv[1]; col = sys.argv[2] if len(sys.argv) > 2 else "stall_no_inst"; top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
cur = None; hdr = None; agg = collections.defaultdict(lambda: [0.0, 0.0, ""])
for r in csv.reader(out.splitlines()):
    if not r: continue
    if r[0] == "File Path": cur = os.path.basename(r[1]); continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or not r[0].isdigit() or r[2] != "-": continue
    d = dict(zip(hdr, r))
    def num(k):
        try: return float(d.get(k, "0") or 0)
        except ValueError: return 0.0
    a = agg[(cur, int(r[0]))]; a[0] += num(col); a[1] += num("# Samples"); a[2] = r[1].strip()[:100]
t = sum(a[0] for a in agg.values()) or 1
print(f"{col}: {t:.0f} samples of {sum(a[1] for a in agg.values()):.0f}")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100 * a[0] / t:5.1f}%  ({a[0]:.0f} of {a[1]:.0f} at this line)  {k[0]}:{k[1]}: {a[2]}")
