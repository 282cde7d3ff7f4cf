#!/usr/bin/env python
"""dft_detect throughput on B200 (BASELINE.json configs[2]): full header bank over N synthetic 48 kHz FM channels.
Prints one JSON line: GPU Msamples/s (device-resident and through the host-buffer C-ABI), per-kernel device times,
HBM roofline of the front end, FP32 rate of the correlator, parity of a channel subset against the oracle, and the
reference dft_detect binary timed on the host cores.   python profiles/dft_bench.py [nchan] [nsamp]"""
import json, os, subprocess, sys, time
from concurrent.futures import ThreadPoolExecutor
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from radiosonde_auto_rx_b200 import dft, siggen
from oracle import pyoracle as po

nchan = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
nsamp = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 18
kinds = list(siggen.HEADER_BANK) + ["noise", "noise", "noise"]
nthreads = max(1, len(os.sched_getaffinity(0)))

def gen(c):
    k = kinds[c % len(kinds)]
    if k == "noise":
        return siggen.noise_iq(9000 + c, nsamp)
    tail = ""
    if k == "M10":      # a real M10/M20 frame continues Manchester-coded (length, type bytes): dft_detect.c:932-965 slices exactly those
        tail = "".join("10" if b == "1" else "01" for b in ("0110010010011111" if (c // 16) % 2 == 0 else "0100010100100000"))
    return siggen.fm_header_iq(9000 + c, nsamp, k, snr_db=(10, 14, 20)[c % 3], invert=(c % 5 == 4), tail=tail)

t0 = time.time()
iq = np.empty((nchan, nsamp, 2), np.int16)
with ThreadPoolExecutor(nthreads) as ex:
    for c, a in enumerate(ex.map(gen, range(nchan))):
        iq[c] = a
t_gen = time.time() - t0
b = dft.DftDetectBatch(48000, 15.0, True, nchan=nchan)
d = torch.from_numpy(iq).cuda()
for _ in range(2):
    res = b.run_device(d, "cs16", nsamp, cont=True)
torch.cuda.synchronize()
t0 = time.perf_counter(); res = b.run_device(d, "cs16", nsamp, cont=True); t_dev = time.perf_counter() - t0
front_ms, corr_ms = b.kernel_ms()
h = torch.from_numpy(iq).pin_memory()
t0 = time.perf_counter(); res2 = b.run(h.numpy(), cont=True); t_host = time.perf_counter() - t0
# parity on a subset
ncheck, bad = min(39, nchan), 0
for c in range(ncheck):
    rc, dets, _ = po.dft_oracle_run(iq[c], 48000, 15.0, True, True)
    g_rc, g = res[c]
    ok = rc == g_rc and min(len(dets), 64) == len(g) and all(x.type == e[0] and x.pos == e[2] and abs(x.score - e[1]) < 2e-4 for x, e in zip(g, dets))
    bad += not ok
    if not ok and bad <= 3:
        sys.stderr.write(f'MISMATCH ch{c} kind={kinds[c % len(kinds)]} rc {rc}/{g_rc} n {len(dets)}/{len(g)}\n')
        for x, e in list(zip(g, dets))[:64]:
            if not (x.type == e[0] and x.pos == e[2] and abs(x.score - e[1]) < 2e-4): sys.stderr.write(f'   got {x.type} {x.score:.5f} {x.pos}  oracle {e[0]} {e[1]:.5f} {e[2]}\n'); break
# reference binary on host threads: each thread runs whole channels through oracle/_ref/dft_detect -c
exe = os.path.join(po.REF_DIR, "dft_detect")
def ref_one(c):
    r = subprocess.run([exe, "-c", "--iq", "--bw", "15", "--dc", "-", "48000", "16"], input=iq[c].tobytes(), capture_output=True)
    return nsamp
nref = min(nchan, 2 * nthreads)
t0 = time.perf_counter()
with ThreadPoolExecutor(nthreads) as ex:
    tot = sum(ex.map(ref_one, range(nref)))
t_ref = time.perf_counter() - t0
total = float(nchan) * nsamp
ndet = sum(len(r[1]) for r in res)
info = b.info
flop = 2.0 * (info.K + 1) * sum([614, 640, 800, 640, 400, 200, 160, 960, 800, 640, 1280, 400, 384, 320]) * (nsamp // (info.K - 4)) * nchan
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
line = {"metric": "dft_detect IQ Msamples/s, full header bank (14 correlators), --iq --bw 15 --dc -c", "nchan": nchan, "nsamp": nsamp,
        "value_device_resident": total / t_dev / 1e6, "value_host_buffers": total / t_host / 1e6, "unit": "Msamples/s",
        "front_ms": front_ms, "corr_ms": corr_ms, "wall_ms_device": t_dev * 1e3, "wall_ms_host": t_host * 1e3,
        "front_roofline": {"bound": "hbm", "achieved_GBs": total * (4 * 2 + 8) / (front_ms * 1e-3) / 1e9, "peak_GBs": peak,
                           "bytes_per_sample": "4 B cs16 read twice (DC mean pass + filter pass) + 8 B FM out"},
        "corr_fp32": {"achieved_TFLOPs": flop / (corr_ms * 1e-3) / 1e12, "note": "direct time-domain matched filter, FP32 FMA (no tensor cores: the reference result is pinned per sample position)"},
        "detections": ndet, "parity_subset": {"channels": ncheck, "mismatches": bad},
        "cpu_reference": {"value": tot / t_ref / 1e6, "unit": "Msamples/s", "threads": nthreads, "sample": f"{nref} channels through oracle/_ref/dft_detect -c"},
        "gen_s": t_gen, "gpu_launches": b.launches}
print(json.dumps(line))
