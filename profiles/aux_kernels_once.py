#!/usr/bin/env python
"""One invocation of every kernel of the widened rows: channeliser + detector front/correlator, frame sync with --ecc2,
periodogram.  Meant to run under `ncu --set full -k regex:...` (profiles/r01_aux_kernels_ncu.json); prints nothing that is a bench value."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from radiosonde_auto_rx_b200 import dft, fsk, siggen
import rs41_frames as rf

sr, nch = 9600000, 128
wide = siggen.wideband_mix(5, sr // 2, sr, [("RS41", 1.2e6), ("DFM9", -2.0e6), ("RS92", 3.1e6)], snr_db=5.0)
d = torch.from_numpy(wide).cuda()
b = dft.DftWidebandBatch(sr, 15.0, True, nchan=nch)
fq = np.linspace(-0.45, 0.45, nch)
res = b.run_device(d, "cs16", wide.shape[0], fq, cont=True)
f, p, step, ms = dft.scan_psd(d, "cs16", wide.shape[0], sr, 0.0, 16384)
streams = [rf.soft_stream(200 + i, 10, (0.30, 0.36, 0.40, 0.43)[i % 4])[0] for i in range(8)]
n = max(s.size for s in streams)
buf = np.zeros((128, n), np.float32)
for c in range(128):
    buf[c, : streams[c % 8].size] = streams[c % 8]
got = fsk.rs41_framesync_device(torch.from_numpy(buf).cuda(), n, n, 128, ecc=2)
print("ran: detections", sum(len(r[1]) for r in res), "frames", len(got), "psd bins", len(p))
