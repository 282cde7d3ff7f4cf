#!/usr/bin/env python
"""Per-role wait/total cycles of fsk_duo_kernel (FSK_PROFILE build, B200FSK_KERNEL=duo). Run on the GPU box."""
import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["B200FSK_KERNEL"] = "duo"
from radiosonde_auto_rx_b200 import fsk, siggen
fsk.LIB_PATH = os.environ.get("B200FSK_LIB", os.path.join(ROOT, "radiosonde_auto_rx_b200", "libb200fsk_prof.so"))
import torch
nchan = int(sys.argv[1]) if len(sys.argv) > 1 else 256
nsamp = int(sys.argv[2]) if len(sys.argv) > 2 else 262144
iq = siggen.fsk_batch(1000, min(nchan, 64), nsamp)
x = torch.from_numpy(np.tile(iq, ((nchan + 63) // 64, 1, 1))[:nchan].astype(np.float32) / np.float32(1000)).cuda()
b = fsk.FskBatch(48000, 4800, 2, 5, 300, 1000, 5000, -5000, 5000, 1, nchan=nchan)
L = b.L
L.b200fsk_prof_read.argtypes = [C.c_void_p, C.c_void_p]
buf = np.zeros((nchan, 16), np.int64)
L.b200fsk_prof_read(b.h, buf.ctypes.data)
maxf = b.max_frames(nsamp)
out = torch.zeros((nchan, maxf * b.Nbits), dtype=torch.float32, device="cuda")
for _ in range(2):
    b.demod_device(x, "f32", nsamp, out, soft=True, in_stride=nsamp, out_stride=maxf * b.Nbits)
torch.cuda.synchronize()
L.b200fsk_prof_read(b.h, buf.ctypes.data)
b.demod_device(x, "f32", nsamp, out, soft=True, in_stride=nsamp, out_stride=maxf * b.Nbits)
torch.cuda.synchronize()
L.b200fsk_prof_read(b.h, buf.ctypes.data)
frames, _ = b.counts()
per = buf.astype(float) / frames[:, None]
m = per.mean(0); m0 = per[0::2].mean(0)
print(f"nchan={nchan} kernel={b.last_kernel()} kernel_ms={b.last_kernel_ms():.3f} => {b.last_kernel_ms()*1e-3*1.965e9/frames.mean():.0f} cycles/frame")
for name, s in (("C0", 0), ("C1", 4), ("E0", 1), ("E1", 2), ("E2", 5), ("D ", 3)):
    print(f"  {name}: wait {m[2*s]:8.0f}  total {m[2*s+1]:8.0f}  busy {m[2*s+1]-m[2*s]:8.0f}")
print(f"  OSC: stepping {m0[12]:8.0f} ({m0[12]/3000:.1f} cyc/step)  events {m0[13]:8.0f}  total {m0[14]:8.0f}")
if os.environ.get("DUO_DETAIL"):
    print(f"  C1 detail: mix {m[4]:8.0f}  integrate {m[5]:8.0f}  publish {m[10]:8.0f}")
    print(f"  D  detail: terms {m[2]:8.0f}  accumulate {m[3]:8.0f}  time+dec {m[11]:8.0f}")
