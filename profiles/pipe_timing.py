#!/usr/bin/env python
"""Per-role wait/busy cycles of fsk_demod_pipe_kernel (FSK_PROFILE build). Run on the GPU box."""
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
L.b200fsk_prof_read(b.h, buf.ctypes.data)
out, frames = b.push(iq, "cs16")
L.b200fsk_prof_read(b.h, buf.ctypes.data)
out, frames = b.push(iq, "cs16")
L.b200fsk_prof_read(b.h, buf.ctypes.data)
per = buf.astype(float) / frames[:, None]
m = per.mean(0)
print(f"nchan={nchan} nsamp={nsamp} frames/ch={frames.mean():.1f} kernel_ms={b.last_kernel_ms():.3f} "
      f"=> {b.last_kernel_ms()*1e-3*1.965e9/frames.mean():.0f} cycles/frame")
print(f"  OSC : wait est {m[0]:8.0f}  wait ring {m[1]:8.0f}  recurrence {m[2]:8.0f} ({m[2]/3000:.1f} cyc/step)  total {m[3]:8.0f}")
print(f"  CONS: wait pub/decide {m[4]:8.0f}  wait osc {m[5]:8.0f}  total {m[6]:8.0f}")
print(f"  CTRL: wait fint {m[7]:8.0f}  total {m[8]:8.0f}")
print(f"  EST : wait publish {m[11]:8.0f}  load {m[12]:8.0f}  fft+mag {m[13]:8.0f}  barriers+IIR {m[14]:8.0f}  total {m[15]:8.0f}")
