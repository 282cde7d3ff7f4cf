#!/usr/bin/env python
"""Static SASS size of one kernel by source region: usage code_size.py <lib.so> <mangled kernel substring> <file.cuh> [bin]
Every instruction is attributed to the outermost line of <file.cuh> it was inlined into; prints instruction counts per `bin` lines.
Used to keep the hot loops of fsk_duo_kernel inside the 32 KB instruction cache level (profiles/r02_notes.md)."""
import re, subprocess, sys, collections, tempfile, os, glob
lib, kern, src = sys.argv[1], sys.argv[2], sys.argv[3]
binw = int(sys.argv[4]) if len(sys.argv) > 4 else 10
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=d, capture_output=True)
for cub in sorted(glob.glob(d + "/*.cubin"), key=os.path.getsize):
    txt = subprocess.run(["nvdisasm", "--print-line-info", cub], capture_output=True, text=True).stdout
    m = re.search(r"^\.text\.(\S*%s\S*):" % re.escape(kern), txt, re.M)
    if m: break
else:
    sys.exit("kernel not found")
name = m.group(1)
body = txt[m.end():]
nxt = re.search(r"^//-+ \.text\.", body, re.M)
if nxt: body = body[:nxt.start()]
last = None; bins = collections.Counter(); n = 0
for l in body.splitlines():
    mm = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', l)
    if mm:
        f1, l1, f2, l2 = mm.group(1), int(mm.group(2)), mm.group(3), mm.group(4)
        if f2 and f2.endswith(src): last = int(l2)
        elif f1.endswith(src): last = l1
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        n += 1; bins[(last // binw) * binw if last else -1] += 1
print(f"{name}: {n} instructions = {n * 16 / 1024:.1f} KB")
for b in sorted(bins): print(f"  {src}:{b:4d}+  {bins[b]:5d}")
