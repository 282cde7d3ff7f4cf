#!/usr/bin/env python
"""Per-phase cycle breakdown of fsk_demod_kernel (FSK_PROFILE build: thread 0 accumulates clock64 deltas).
Run on the GPU box:  python profiles/phase_timing.py [nchan] [nsamp]"""
import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from radiosonde_auto_rx_b200 import fsk, siggen
fsk.LIB_PATH = os.path.join(ROOT, "radiosonde_auto_rx_b200", "libb200fsk_prof.so")
nchan = int(sys.argv[1]) if len(sys.argv) > 1 else 256
nsamp = int(sys.argv[2]) if len(sys.argv) > 2 else 262144
iq = siggen.fsk_batch(1000, nchan, nsamp)
b = fsk.FskBatch(48000, 4800, 2, 5, 300, 1000, 5000, -5000, 5000, 1, nchan=nchan)
L = b.L
L.b200fsk_prof_read.argtypes = [C.c_void_p, C.c_void_p]
buf = np.zeros((nchan, 16), np.int64)
L.b200fsk_prof_read(b.h, buf.ctypes.data)            # attach
out, frames = b.push(iq, "cs16")                        # warm
L.b200fsk_prof_read(b.h, buf.ctypes.data)
out, frames = b.push(iq, "cs16")
L.b200fsk_prof_read(b.h, buf.ctypes.data)
names = ["load", "A fft+Sf", "B estimator", "C osc(+shift)", "C mix", "D integrate", "E terms", "E chain", "E atan2/nin",
         "F decide", "F ebno+rec", "-"]
per = buf[:, :12].astype(float) / frames[:, None]
tot = per.sum(1).mean()
print(f"nchan={nchan} nsamp={nsamp} frames/ch={frames.mean():.1f} kernel_ms={b.last_kernel_ms():.3f} cycles/frame={tot:.0f}")
for k, n in enumerate(names[:11]):
    print(f"  {n:16s} {per[:, k].mean():9.0f} cyc  {100 * per[:, k].mean() / tot:5.1f}%")
