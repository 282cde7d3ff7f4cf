#!/usr/bin/env python
"""Stall reasons per source REGION of one kernel, inlined callees included: ncu_roles.py <rep> <lib.so> <kernel substring> <file.cuh> <line:name> ...
Every SASS instruction is attributed to the outermost line of <file.cuh> it was inlined into (nvdisasm line info of the library the
capture was taken with), the regions are given as first-line:name pairs in ascending order."""
import csv, re, subprocess, sys, collections, tempfile, os, glob
rep, lib, kern, src = sys.argv[1:5]
regs = [(int(a.split(":")[0]), a.split(":")[1]) for a in sys.argv[5:]]
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=d, capture_output=True)
body = None
for cub in sorted(glob.glob(d + "/*.cubin"), key=os.path.getsize):
    txt = subprocess.run(["nvdisasm", "--print-line-info", cub], capture_output=True, text=True).stdout
    m = re.search(r"^\.text\.(\S*%s\S*):" % re.escape(kern), txt, re.M)
    if m:
        body = txt[m.end():]
        nxt = re.search(r"^//-+ \.text\.", body, re.M)
        if nxt: body = body[:nxt.start()]
        break
if body is None: sys.exit("kernel not found")
lines = []; last = None
for l in body.splitlines():
    mm = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', l)
    if mm:
        f1, l1, f2, l2 = mm.group(1), int(mm.group(2)), mm.group(3), mm.group(4)
        if f2 and f2.endswith(src): last = int(l2)
        elif f1.endswith(src): last = l1
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l): lines.append(last)
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
h = rows[hi]; R = [r for r in rows[hi + 1:] if r]
if len(R) != len(lines): sys.exit(f"instruction counts differ: capture {len(R)} vs library {len(lines)} (different build?)")
ix = {k: i for i, k in enumerate(h)}
st = [k for k in h if k.startswith("stall_") and "Not Issued" not in k]
def f(x):
    try: return float(x or 0)
    except ValueError: return 0.0
def region(n):
    name = "prologue"
    for a, nm in regs:
        if n is not None and n >= a: name = nm
    return name
agg = collections.defaultdict(collections.Counter); smp = collections.Counter(); ins = collections.Counter()
for r, n in zip(R, lines):
    g = region(n); smp[g] += f(r[ix["# Samples"]]); ins[g] += f(r[ix["Instructions Executed"]])
    for k in st: agg[g][k[6:]] += f(r[ix[k]])
tot = sum(smp.values()) or 1
for g in smp:
    print(f"{g:10s} {100 * smp[g] / tot:5.1f}% of samples  " + "  ".join(f"{k} {100 * v / max(smp[g], 1):.0f}%" for k, v in agg[g].most_common(7)))
