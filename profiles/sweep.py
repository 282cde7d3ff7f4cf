#!/usr/bin/env python
"""Throughput of the batched fsk_demod kernels vs channel count (device-resident float2 input, RS41 production line).
usage: sweep.py [nsamp_log2] mode:nchan [mode:nchan ...]     mode = auto | pipe | wpc | duo | serial
Channels beyond 256 repeat the same 256 seeded streams (throughput only; parity is tests/)."""
import os, sys, json, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from radiosonde_auto_rx_b200 import fsk, siggen

lg = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 18
specs = [a for a in sys.argv[1:] if ":" in a] or ["pipe:256", "wpc:256", "wpc:1024", "wpc:4096"]
nsamp = 1 << lg
base = siggen.fsk_batch(1000, 256, nsamp, workers=max(1, len(os.sched_getaffinity(0))))
xb = torch.from_numpy(base.astype(np.float32) / np.float32(1000)).cuda()
res = []
for sp in specs:
    mode, nchan = sp.split(":"); nchan = int(nchan)
    os.environ.pop("B200FSK_KERNEL", None)
    if mode != "auto":
        os.environ["B200FSK_KERNEL"] = mode
    reps = (nchan + 255) // 256
    d_in = xb.repeat(reps, 1, 1)[:nchan].contiguous()
    b = fsk.FskBatch(48000, 4800, 2, 5, 300, 1000, 5000, -5000, 5000, 1, nchan=nchan, device=0)
    maxf = b.max_frames(nsamp)
    d_out = torch.zeros((nchan, maxf * b.Nbits), dtype=torch.float32, device="cuda")
    st = torch.cuda.Stream()
    for _ in range(2):
        b.demod_device(d_in, "f32", nsamp, d_out, soft=True, in_stride=nsamp, out_stride=maxf * b.Nbits, stream=st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    steps = 3
    e0.record(st)
    for _ in range(steps):
        b.demod_device(d_in, "f32", nsamp, d_out, soft=True, in_stride=nsamp, out_stride=maxf * b.Nbits, stream=st)
    e1.record(st)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    frames, consumed = b.counts()
    gsps = float(consumed.sum()) / (ms * 1e-3) / 1e9
    r = {"mode": mode, "nchan": nchan, "nsamp": nsamp, "ms": ms, "gsps": gsps, "frames": int(frames[0])}
    res.append(r)
    print(json.dumps(r), flush=True)
    del b, d_in, d_out
    torch.cuda.empty_cache()
