#!/usr/bin/env python
"""Wideband front on B200: ONE baseband capture, nchan candidate frequencies, each translated + decimated to the
IF rate (dft_detect --IQ <fq>, dft_detect.c:751-761) and run through the full detector.  Prints one JSON line: device times of
the channeliser and of the detector behind it, channel-samples/s, FP32 rate of the decimating FIR, parity of a subset against
the oracle, and the reference binary (`dft_detect -c --IQ fq ...`, one process per candidate) timed on the host threads.
    python profiles/wideband_bench.py [sr_base] [nchan] [seconds]"""
import json, os, subprocess, sys, time
from concurrent.futures import ThreadPoolExecutor
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from radiosonde_auto_rx_b200 import dft, siggen
from oracle import pyoracle as po

sr = int(sys.argv[1]) if len(sys.argv) > 1 else 9600000
nchan = int(sys.argv[2]) if len(sys.argv) > 2 else 256
secs = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
nthreads = max(1, len(os.sched_getaffinity(0)))
nwide = int(sr * secs)
kinds = ["RS41", "DFM9", "RS92", "LMS6", "IMET5", "MRZ", "MEISEI", "RD94RD41"]
rng = np.random.default_rng(4)
grid = np.linspace(-0.45, 0.45, nchan) * sr                      # candidate centre frequencies, Hz
grid = np.rint(grid / 1000.0) * 1000.0
occupied = sorted(rng.choice(nchan, size=min(8, nchan), replace=False).tolist())
emit = [(kinds[i % len(kinds)], float(grid[c])) for i, c in enumerate(occupied)]
t0 = time.time()
wide = siggen.wideband_mix(77, nwide, sr, emit, snr_db=3.0, dc=(150.0, -90.0))
t_gen = time.time() - t0
fqs = grid / sr
b = dft.DftWidebandBatch(sr, 15.0, True, nchan=nchan)
d = torch.from_numpy(wide).cuda()
for _ in range(2):
    res = b.run_device(d, "cs16", nwide, fqs, cont=True)
torch.cuda.synchronize()
t0 = time.perf_counter(); res = b.run_device(d, "cs16", nwide, fqs, cont=True); t_dev = time.perf_counter() - t0
chan_ms = b.channelise_ms(); front_ms, corr_ms = b.kernel_ms()
hp = torch.from_numpy(wide).pin_memory()
t0 = time.perf_counter(); res2 = b.run(hp.numpy(), fqs, cont=True); t_host = time.perf_counter() - t0
# parity: every occupied candidate + a few empty ones against the oracle
check = occupied + [c for c in range(0, nchan, max(1, nchan // 8)) if c not in occupied][:8]
bad = 0
for c in check:
    rc, dets, _, _ = po.dft_oracle_run_wideband(wide, sr, float(fqs[c]), 15.0, True, True)
    g_rc, g = res[c]
    ok = rc == g_rc and len(dets) == len(g) and all(x.type == e[0] and x.pos == e[2] and abs(x.score - e[1]) < 2e-4 for x, e in zip(g, dets))
    ok = ok and res2[c][0] == g_rc and len(res2[c][1]) == len(g)
    bad += not ok
    if not ok:
        sys.stderr.write(f"MISMATCH cand {c} {grid[c]:+.0f} Hz: oracle {rc} {dets[:2]}  gpu {g_rc} {g[:2]}\n")
exe = os.path.join(po.REF_DIR, "dft_detect")
raw = wide.tobytes()
def ref_one(c):
    subprocess.run([exe, "-c", "--IQ", repr(float(fqs[c])), "--bw", "15", "--dc", "-", str(sr), "16"], input=raw, capture_output=True)
    return nwide
nref = min(nchan, nthreads)
t0 = time.perf_counter()
with ThreadPoolExecutor(nthreads) as ex:
    tot = sum(ex.map(ref_one, range(nref)))
t_ref = time.perf_counter() - t0
nout = nwide // b.decM
total = float(nchan) * nwide                                      # channel-samples: every candidate consumes the whole capture
line = {"metric": "wideband dft_detect --IQ: candidate-channel x wide-sample throughput", "sr_base": sr, "if_sr": b.sample_rate, "decM": b.decM,
        "dectaps": b.dectaps, "nchan": nchan, "nwide": nwide, "occupied": len(emit),
        "value_device_resident": total / t_dev / 1e6, "value_host_buffers": total / t_host / 1e6, "unit": "M channel-samples/s",
        "realtime_factor_device": secs / t_dev, "channelise_ms": chan_ms, "detect_front_ms": front_ms, "detect_corr_ms": corr_ms,
        "wall_ms_device": t_dev * 1e3, "wall_ms_host": t_host * 1e3,
        "channeliser": {"rotations_per_s": total / (chan_ms * 1e-3), "fir_TFLOPs": 4.0 * b.dectaps * nout * nchan / (chan_ms * 1e-3) / 1e12,
                        "bytes_algorithmic": nwide * 4 + nchan * nout * 8,
                        "note": "one FP64 sincos + complex multiply per (candidate, wide sample); FIR 2 FMA per tap per output; the capture is read once "
                                "from HBM and then from L2 by the other candidates"},
        "detections": sum(len(r[1]) for r in res), "parity_subset": {"candidates": len(check), "mismatches": bad},
        "cpu_reference": {"value": tot / t_ref / 1e6, "unit": "M channel-samples/s", "threads": nthreads,
                          "sample": f"{nref} candidates, one oracle/_ref/dft_detect -c --IQ process each"},
        "gen_s": t_gen, "gpu_launches": b.launches}
print(json.dumps(line))
