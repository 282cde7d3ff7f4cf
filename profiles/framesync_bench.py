#!/usr/bin/env python
"""SURVEY 8f-1 + 8f-4 on B200: header sync + frame assembly + Reed-Solomon (--ecc2) + CRC over soft bits resident in HBM, against
`rs41mod -r --ecc2 --softin -i` (one process per channel on the host threads).  One JSON line.
    python profiles/framesync_bench.py [nchan] [frames_per_channel]"""
import json, os, subprocess, sys, time
from concurrent.futures import ThreadPoolExecutor
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from radiosonde_auto_rx_b200 import fsk
from oracle import pyoracle as po
import rs41_frames as rf

nchan = int(sys.argv[1]) if len(sys.argv) > 1 else 256
nfr = int(sys.argv[2]) if len(sys.argv) > 2 else 20
nthreads = max(1, len(os.sched_getaffinity(0)))
sig = (0.30, 0.36, 0.40, 0.43)
uniq = [rf.soft_stream(100 + i, nfr, sig[i % 4])[0] for i in range(16)]            # 16 distinct streams, tiled over the channels
n = max(s.size for s in uniq)
buf = np.zeros((nchan, n), np.float32)
for c in range(nchan):
    s = uniq[c % 16]; buf[c, : s.size] = s
nbits = np.array([uniq[c % 16].size for c in range(nchan)], np.int32)
d = torch.from_numpy(buf).cuda(); nb = torch.from_numpy(nbits).cuda()
res = {}
for ecc in (0, 2):
    for _ in range(2):
        got = fsk.rs41_framesync_device(d, n, nb, nchan, ecc=ecc)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    got = fsk.rs41_framesync_device(d, n, nb, nchan, ecc=ecc)
    res[ecc] = (time.perf_counter() - t0, got, fsk.rs41_framesync_kernel_ms())
t2, got, k2 = res[2]
bad = 0
for c in range(16):
    exp = po.softsync_oracle(uniq[c], False, True, ecc=2)
    mine = [g for g in got if g[0] == c]
    bad += len(mine) != len(exp) or any(g[4].tobytes() != e[3].tobytes() or g[6] != e[5] or g[7] != e[6] for g, e in zip(mine, exp))
exe = os.path.join(po.REF_DIR, "rs41mod")
def ref_one(c):
    subprocess.run([exe, "-r", "--ecc2", "--softin", "-i"], input=uniq[c % 16].tobytes(), capture_output=True); return int(nbits[c])
nref = min(nchan, 2 * nthreads)
t0 = time.perf_counter()
with ThreadPoolExecutor(nthreads) as ex:
    tot = sum(ex.map(ref_one, range(nref)))
t_ref = time.perf_counter() - t0
bits = float(nbits.sum())
print(json.dumps({"metric": "soft bits/s through header sync + frame assembly (+ RS(255,231) x2 --ecc2 + CRC)", "nchan": nchan, "bits_per_channel": int(n),
                  "frames": len(got), "frames_ok": sum(1 for g in got if g[6] >= 0), "frames_corrected": sum(1 for g in got if g[6] > 0),
                  "value_sync_only_Mbit_s": bits / res[0][0] / 1e6, "value_sync_ecc2_Mbit_s": bits / t2 / 1e6, "ms_sync_only": res[0][0] * 1e3, "ms_sync_ecc2": t2 * 1e3, "kernel_ms_sync_only": res[0][2], "kernel_ms_sync_ecc2": k2,
                  "note": "wall time of the C-ABI call: kernel + frame records D2H (2.6 KB per frame)", "parity_channels": 16, "mismatches": int(bad),
                  "cpu_reference": {"value_Mbit_s": tot / t_ref / 1e6, "threads": nthreads, "sample": f"{nref} channels, one oracle/_ref/rs41mod -r --ecc2 --softin -i each"}}))
