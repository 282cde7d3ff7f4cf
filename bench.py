#!/usr/bin/env python
"""bench.py -- IQ Msamples/s through the batched fsk_demod path (BASELINE.json metric).

Workload (BASELINE.json configs[1]): 256 concurrent 48 kHz 2FSK (RS41 production line:
``fsk_demod --cs16 -b -5000 -u 5000 -s --mask 5000 --nsym=300 -p 5 2 48000 4800``,
auto_rx/autorx/decode.py:901) channels x 2^20 samples each, synthetic and seeded (one of them noise only).

One *step* = one pass of the hot path over that batch (one kernel launch: every channel walks its
2^20-sample block frame by frame).  Reported on ONE JSON line:
  value     whole-job Msamples/s with the IQ already resident in HBM as float2 (8 B/sample read once)
  e2e       the same metric through the host-buffer C-ABI call (b200fsk_push_host): pinned cs16 host
            buffers, H2D + kernel + D2H of the soft bits inside the timed region
  parity    BASELINE.md 3.5 gate: the timed outputs (device-resident AND end-to-end) of >= 16 channels are compared bit for
            bit with the reference modem (oracle/_ref/libfsk_ref.so = the unmodified utils/fsk.c) -- on a mismatch NO value
            is printed and the exit code is 1
  roofline  achieved algorithmic HBM GB/s of the frame kernel vs the measured copy peak
  config.sweep   device-resident throughput at 256 / 1024 / 4096 channels per GPU (the larger batches switch to the
            many-channels-per-CTA kernel), each parity-checked on a channel subset
  cpu_baseline   the reference's own CPU implementation timed on this box's host cores: one pinned process per logical
            CPU running the reference frame pump over the SAME channels, median of 3 passes, plus a 1-core figure
`--impl reference` times the reference CPU implementation alone (rank 0 only under torchrun).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "IQ Msamples/s through batched fsk_demod (bit-exact output)"
UNIT = "Msamples/s"
FS, RS, M, P, NSYM, MASK, LO, HI = 48000, 4800, 2, 5, 300, 5000, -5000, 5000
BYTES_PER_SAMPLE_F32 = 8.0 + 4.0 / 10.0          # float2 in + one float soft bit per Ts=10 samples
NOISE_CHANNEL = 5                                # this channel of every rank's batch carries noise only (nin changes often)


def workload_name(nchan, nsamp):
    return f"batched fsk_demod RS41-prod 2FSK 48000/4800 P5 nsym300 mask5000: {nchan} ch x {nsamp} samples/GPU"


def gen_channels(rank: int, nchan: int, nsamp: int) -> np.ndarray:
    from radiosonde_auto_rx_b200 import siggen
    iq = siggen.fsk_batch(1000 + 100000 * rank, nchan, nsamp, workers=max(1, len(os.sched_getaffinity(0))))
    if nchan > NOISE_CHANNEL:
        iq[NOISE_CHANNEL] = siggen.noise_iq(424242 + rank, nsamp)
    return iq


def cpu_quota() -> float | None:
    """CPUs this container may really use (cgroup CFS quota), or None when unlimited: the affinity mask of a GPU box lists every
    logical CPU of the host although the container is throttled to a few of them -- timing 128 processes then measures the throttle."""
    try:
        q, per = open("/sys/fs/cgroup/cpu.max").read().split()
        if q != "max":
            return float(q) / float(per)
    except (OSError, ValueError):
        pass
    try:
        q = float(open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read())
        per = float(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
        if q > 0:
            return q / per
    except (OSError, ValueError):
        pass
    return None


def usable_cpus() -> list[int]:
    cpus = sorted(os.sched_getaffinity(0))
    q = cpu_quota()
    if q is not None:
        cpus = cpus[: max(1, int(q))]
    return cpus


def bind_to_gpu_numa(local: int) -> str:
    """Run this process (and allocate its pinned buffers) on the CPUs NVML names as local to the GPU (its NUMA node)."""
    try:
        import pynvml
        import torch
        pynvml.nvmlInit()
        h, idx = None, local
        try:                                              # the CUDA device's own identity: immune to CUDA_VISIBLE_DEVICES remapping
            uuid = str(torch.cuda.get_device_properties(local).uuid)
            h = pynvml.nvmlDeviceGetHandleByUUID(uuid if uuid.startswith("GPU-") else "GPU-" + uuid)
            idx = pynvml.nvmlDeviceGetIndex(h)
        except Exception:                                 # noqa: BLE001
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[local]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else local
            h = pynvml.nvmlDeviceGetHandleByIndex(idx)
        before = len(os.sched_getaffinity(0))
        pynvml.nvmlDeviceSetCpuAffinity(h)
        after = sorted(os.sched_getaffinity(0))
        return f"bound to the {len(after)} CPUs local to GPU {idx} (of {before}): {after[0]}..{after[-1]}"
    except Exception as e:                                # noqa: BLE001
        return f"not NUMA-bound ({type(e).__name__}: {e})"


def cpu_model() -> str:
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.startswith("model name"):
                return ln.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


# ---------------------------------------------------------------------------------------------------
# reference arm: the UNMODIFIED reference modem (oracle/_ref/libfsk_ref.so) on the host cores
# ---------------------------------------------------------------------------------------------------
def baseline_kind():
    from oracle import pyoracle as po
    if os.path.exists(os.path.join(po.REF_DIR, "libfsk_ref.so")) or (os.path.isdir("/root/reference/utils") and not os.environ.get("B200_BENCH_NO_REF")):
        return "reference", "oracle/_ref/libfsk_ref.so (unmodified utils/fsk.c + kiss_fft, reference flags)"
    return "port", "oracle/liboracle.so (C restatement; oracle/_ref absent)"


def _ref_pump(iq_row: np.ndarray, want_out: bool, passes: int = 1):
    """Reference frame pump (utils/fsk_demod.c:281-436 over memory) of one channel; no per-frame statistics.
    passes > 1: the same modem object walks the buffer that many times (state carried over, like a library handle that is
    launched repeatedly on one resident buffer); the output is that of the last pass."""
    from oracle import pyoracle as po
    import ctypes as C
    cfg = po.RS41
    if baseline_kind()[0] == "reference":
        L = po.reflib()
        nsamp = iq_row.shape[0]
        nb = cfg.Nsym
        maxf = nsamp // (cfg.Nsym * (cfg.Fs // cfg.Rs) - 5) + 2
        out = np.zeros(maxf * nb, np.float32)
        cons = C.c_long(0)
        nf = L.ref_fsk_run_stream_passes(cfg.Fs, cfg.Rs, cfg.M, cfg.P, cfg.Nsym, cfg.fs_tx, cfg.lo, cfg.hi, 1, iq_row.ctypes.data, nsamp, 0, 1, 0,
                                         out.ctypes.data, C.byref(cons), None, 0, None, passes)
        return (out[: nf * nb] if want_out else None), cons.value
    orc = po.Oracle(cfg)
    for _ in range(passes):
        out, cons, _ = orc.run_stream(iq_row, "cs16")
    return (out if want_out else None), cons


def _worker(core, chans, iq, start_evt, q, repeats, bar=None):
    try:
        os.sched_setaffinity(0, {core})
    except OSError:
        pass
    _ref_pump(iq[chans[0], :30000], False)            # load the library, touch the code
    q.put(("ready", core))
    start_evt.wait()
    res = []
    for _ in range(repeats):
        if bar is not None:
            bar.wait()                                 # every pass starts on all processes together: a pass's wall time is one number
        t0 = time.perf_counter()
        n = 0
        for c in chans:
            n += _ref_pump(iq[c], False)[1]
        res.append((n, time.perf_counter() - t0, t0))
    q.put(("done", core, res))


def time_reference(iq: np.ndarray, repeats: int = 3, ncores: int | None = None):
    """One pinned process per logical CPU, channel c -> process c mod P, `repeats` passes over all channels.
    Returns dict(value = median Msamples/s over the passes (whole-batch samples / slowest process), ...)."""
    import multiprocessing as mp
    cpus = usable_cpus()
    if ncores:
        cpus = cpus[:ncores]
    nchan = iq.shape[0]
    nproc = min(len(cpus), nchan)
    ctx = mp.get_context("fork")
    q, start = ctx.Queue(), ctx.Event()
    bar = ctx.Barrier(nproc)
    procs = []
    for r in range(nproc):
        chans = list(range(r, nchan, nproc))
        pr = ctx.Process(target=_worker, args=(cpus[r], chans, iq, start, q, repeats, bar))
        pr.start()
        procs.append(pr)
    ready = 0
    while ready < nproc:
        if q.get()[0] == "ready":
            ready += 1
    start.set()
    results = []
    while len(results) < nproc:
        m = q.get()
        if m[0] == "done":
            results.append(m[2])
    for pr in procs:
        pr.join()
    per_pass = []
    for k in range(repeats):
        n = sum(r[k][0] for r in results)
        t_begin = min(r[k][2] for r in results)
        t_end = max(r[k][2] + r[k][1] for r in results)
        per_pass.append(n / (t_end - t_begin) / 1e6)
    per_core = [r[k][0] / r[k][1] / 1e6 for r in results for k in range(repeats)]
    return {"value": float(np.median(per_pass)), "passes": [float(x) for x in per_pass], "cores": nproc,
            "logical_cpus": len(os.sched_getaffinity(0)), "cpu_quota": cpu_quota(),
            "per_core_median": float(np.median(per_core)), "samples_per_pass": int(sum(r[0][0] for r in results))}


def one_core_reference(iq: np.ndarray, nchan: int = 4):
    return time_reference(iq[:nchan], repeats=3, ncores=1)["value"]


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    nchan, nsamp = args.channels, args.nsamp
    iq = gen_channels(0, nchan, nsamp)
    kind, what = baseline_kind()
    for _ in range(min(args.warmup, 1)):
        time_reference(iq, repeats=1)
    steps = max(3, args.steps)
    r = time_reference(iq, repeats=steps)
    one = one_core_reference(iq)
    val = r["value"]
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": min(args.warmup, 1), "ms_per_step": r["samples_per_pass"] / (val * 1e6) * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(nchan, nsamp),
                   "reference_step": f"the same {nchan} ch x {nsamp} samples per step; one pinned process per usable CPU ({r['cores']}), "
                                     "reference frame pump without per-frame statistics; value = median over the steps"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": r["cores"], "kind": kind, "cpu_model": cpu_model(),
                         "passes": r["passes"], "per_core_median": r["per_core_median"], "one_core": one,
                         "logical_cpus": r["logical_cpus"], "cpu_quota": r["cpu_quota"],
                         "sample": f"{nchan} ch x {nsamp} samples x {steps} passes, {what}"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------
# clocks sampler
# ---------------------------------------------------------------------------------------------------
class Clocks:
    """nvidia-smi sampler (B200_PROFILING.md clocks line), started before the warm-up so that it is already producing
    samples when the timed region begins; only samples whose timestamp falls inside [t0, t1] are reported."""
    Q = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu: int):
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.p = None

    def stop(self, t0: float, t1: float):
        import datetime
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.1)
        self.p.terminate()
        try:
            out, _ = self.p.communicate(timeout=5)
        except Exception:
            self.p.kill()
            out = ""
        sm_in, sm_all, mx, reasons = [], [], None, set()
        for ln in out.strip().splitlines():
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 10:
                continue
            try:
                ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                v = float(f[2]); mx = float(f[3])
            except ValueError:
                continue
            sm_all.append(v)
            if t0 - 0.02 <= ts <= t1 + 0.02:
                sm_in.append(v)
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[6:10]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
        use = sm_in if sm_in else sm_all
        return {"sm_mhz": float(np.median(use)) if use else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm_in), "samples_total": len(sm_all)}


def committed_traffic(kernel: str, nchan: int, nsamp: int):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of this workload from the committed ncu capture (profiles/traffic.json);
    the run itself cannot measure DRAM traffic, so the figure is labelled with its source."""
    try:
        for t in json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))["captures"]:
            if t.get("kernel") == kernel and t.get("channels") == nchan and t.get("nsamp") == nsamp:
                return float(t["dram_bytes_per_launch"]), t.get("source", "profiles/")
    except Exception:
        pass
    return None, None


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured)"
        except Exception:
            pass
    return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"


def reference_outputs(iq: np.ndarray, chans, passes: int = 1):
    """soft bits of the reference modem for the given channels (threads: the C pump releases the GIL)"""
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=min(len(chans), max(1, len(os.sched_getaffinity(0))))) as ex:
        return dict(zip(chans, ex.map(lambda c: _ref_pump(iq[c], True, passes)[0], chans)))


def count_mismatches(ref: dict, got_rows, frames, nbits: int) -> int:
    bad = 0
    for c, r in ref.items():
        g = np.ascontiguousarray(got_rows(c))[: int(frames[c]) * nbits]
        if int(frames[c]) * nbits != r.size or g.view(np.uint32).tobytes() != r.view(np.uint32).tobytes():
            bad += 1
    return bad


# ---------------------------------------------------------------------------------------------------
# secondary workload: dft_detect, BASELINE.json configs[2] (full header bank over 4096 synthetic 48 kHz FM channels)
# ---------------------------------------------------------------------------------------------------
def run_dft_secondary(args, local: int):
    """One JSON object: device-resident and end-to-end Msamples/s of the batched detector, per-kernel times, honest rooflines (front end
    vs HBM, correlator vs the SIMT FP32 peak), parity of a channel subset against the REFERENCE BINARY's printed detections, and that
    binary timed on the usable host CPUs (one pinned process each)."""
    import torch
    from concurrent.futures import ThreadPoolExecutor
    from oracle import pyoracle as po
    from radiosonde_auto_rx_b200 import dft, siggen
    nchan, nsamp, ndist = args.dft_channels, args.dft_nsamp, 64
    kinds = list(siggen.HEADER_BANK) + ["noise", "noise", "noise"]

    def gen(c):
        k = kinds[c % len(kinds)]
        if k == "noise":
            return siggen.noise_iq(9000 + c, nsamp)
        tail = ""
        if k == "M10":
            tail = "".join("10" if b == "1" else "01" for b in ("0110010010011111" if (c // 16) % 2 == 0 else "0100010100100000"))
        return siggen.fm_header_iq(9000 + c, nsamp, k, snr_db=(10, 14, 20)[c % 3], invert=(c % 5 == 4), tail=tail)
    with ThreadPoolExecutor(len(usable_cpus())) as ex:
        base = np.stack(list(ex.map(gen, range(ndist))))
    # CPU arm + parity reference first (subprocesses of the reference binary)
    exe = os.path.join(po.REF_DIR, "dft_detect")
    have_ref = os.path.exists(exe)
    cpus = usable_cpus()
    cpu = None
    ref_dets = {}
    if have_ref:
        def ref_one(c):
            t0 = time.perf_counter()
            rc, dets = po.ref_cli_dft_detect(base[c], 48000, 15.0, True, True)
            return c, rc, dets, time.perf_counter() - t0
        nref = min(ndist, 2 * len(cpus))
        t0 = time.perf_counter()
        with ThreadPoolExecutor(len(cpus)) as ex:
            rr = list(ex.map(ref_one, range(nref)))
        dt = time.perf_counter() - t0
        for c, rc, dets, _ in rr:
            ref_dets[c] = (rc, dets)
        cpu = {"value": nref * nsamp / dt / 1e6, "unit": UNIT, "cores": len(cpus), "kind": "reference", "cpu_model": cpu_model(),
               "per_process": float(np.median([nsamp / r[3] / 1e6 for r in rr])),
               "sample": f"{nref} channels x {nsamp} samples through oracle/_ref/dft_detect -c -v --iq --bw 15 --dc - 48000 16, one process per usable CPU"}
    reps = (nchan + ndist - 1) // ndist
    d = torch.from_numpy(base).cuda().repeat(reps, 1, 1)[:nchan].contiguous()
    b = dft.DftDetectBatch(48000, 15.0, True, nchan=nchan, device=local)
    for _ in range(2):
        res = b.run_device(d, "cs16", nsamp, cont=True)
    torch.cuda.synchronize()
    steps = 3
    t0 = time.perf_counter()
    for _ in range(steps):
        res = b.run_device(d, "cs16", nsamp, cont=True)
    torch.cuda.synchronize()
    t_dev = (time.perf_counter() - t0) / steps
    front_ms, corr_ms = b.kernel_ms()
    launches = b.launches
    bad = 0
    for c, (rc, dets) in ref_dets.items():
        cc = (nchan - 1) // ndist * ndist + c if (nchan - 1) // ndist * ndist + c < nchan else c     # a replica far into the batch
        g_rc, g = res[cc]
        ok = rc == g_rc and min(len(dets), 64) == len(g) and all(x.type == e[0] and x.pos == e[2] and abs(x.score - e[1]) < 2.5e-4 for x, e in zip(g, dets))
        bad += not ok
    # end to end from pinned host buffers (bounded: 1024 channels)
    ne = min(nchan, 1024)
    be = dft.DftDetectBatch(48000, 15.0, True, nchan=ne, device=local)
    h = torch.from_numpy(np.tile(base, ((ne + ndist - 1) // ndist, 1, 1))[:ne]).pin_memory()
    be.run(h.numpy(), cont=True)
    t0 = time.perf_counter()
    be.run(h.numpy(), cont=True)
    t_host = time.perf_counter() - t0
    info = b.info
    flop = 2.0 * (info.K + 1) * sum([614, 640, 800, 640, 400, 200, 160, 960, 800, 640, 1280, 400, 384, 320]) * (nsamp // (info.K - 4)) * nchan
    peak, peak_src = measured_peak()
    total = float(nchan) * nsamp
    simt_peak = 148 * 128 * 2 * 1.965e9 / 1e12
    out = {"metric": "dft_detect IQ Msamples/s, full header bank (14 correlators), --iq --bw 15 --dc -c  (BASELINE.json configs[2])",
           "workload": f"{nchan} ch x {nsamp} cs16 samples (64 distinct seeded channels, each header type + noise, repeated)",
           "value": total / t_dev / 1e6, "unit": UNIT, "ms_per_step": t_dev * 1e3, "front_ms": front_ms, "corr_ms": corr_ms,
           "e2e": {"value": float(ne) * nsamp / t_host / 1e6, "unit": UNIT, "channels": ne, "h2d_bytes_per_step": float(ne) * nsamp * 4,
                   "api": "b200dft_run_host: pinned cs16 host buffer -> H2D -> kernels -> result records"},
           "roofline_front": {"bound": "hbm", "achieved": total * (4 * 2 + 8) / (front_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                              "frac": total * (4 * 2 + 8) / (front_ms * 1e-3) / 1e9 / peak, "peak_source": peak_src,
                              "bytes_per_sample": "cs16 read twice (DC mean pass + filter pass) + two float FM streams written"},
           "roofline_corr": {"bound": "fp32 SIMT (no dense contraction pinned per sample position: tensor cores not applicable)",
                             "achieved": flop / (corr_ms * 1e-3) / 1e12, "peak": simt_peak, "unit": "TFLOP/s", "frac": flop / (corr_ms * 1e-3) / 1e12 / simt_peak,
                             "peak_source": "148 SM x 128 FP32 lanes x 2 flop x 1.965 GHz (nominal)",
                             "note": "direct time-domain matched filter: ~16.5 kflop per input sample vs ~4 kflop in the reference's FFT formulation"},
           "parity": {"channels": len(ref_dets), "mismatches": bad, "checked_against": "oracle/_ref/dft_detect (reference binary): type, sample position and exit code exact, score within 2e-4 of the printed %.4f value"},
           "gpu_launches": int(launches)}
    if cpu:
        out["cpu_baseline"] = cpu
    return out


# ---------------------------------------------------------------------------------------------------
# --scatter: the collective path (BASELINE.json configs[3]: channels sharded across the GPUs of one box via NCCL scatter)
# ---------------------------------------------------------------------------------------------------
def run_scatter(args):
    """Rank 0 holds the whole capture (world x C channels, cs16, the production sample format) in ITS HBM -- as if one GPU had received
    it -- and scatters it by channel block, one time block after the other (NCCL over NVLink); every rank demodulates block t while
    block t+1 is in flight (two receive buffers, copy stream + compute stream).  One step = T time blocks.  Reports whole-job
    Msamples/s (max over ranks of CUDA-event time) and the scatter rate out of rank 0 against the measured NVLink peer bandwidth.
    Block ends drop the sub-frame remainder like every b200fsk_demod_device call (the modem state carries over)."""
    import torch
    import torch.distributed as dist
    from radiosonde_auto_rx_b200 import fsk
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    numa = bind_to_gpu_numa(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    C, B, T = args.scatter_channels, args.scatter_block, args.scatter_blocks
    dev = torch.device("cuda", local)
    cap = None
    if rank == 0:                                   # the capture: world*C channels x T*B samples, generated once (64 distinct channels, repeated)
        base = torch.from_numpy(gen_channels(0, 64, T * B)).to(dev)
        cap = base.repeat((world * C + 63) // 64, 1, 1)[: world * C].contiguous()          # [world*C, T*B, 2] int16
        del base
    b = fsk.FskBatch(FS, RS, M, P, NSYM, 1000, MASK, LO, HI, 1, nchan=C, device=local)
    maxf = b.max_frames(B)
    d_out = torch.zeros((C, maxf * b.Nbits), dtype=torch.float32, device=dev)
    from radiosonde_auto_rx_b200 import shard
    bs = shard.BlockScatter(C, B, dev, src=0)       # two receive buffers, copy stream + compute stream (radiosonde_auto_rx_b200/shard.py)
    comp_s = bs.comp_stream

    def consume(buf, t, stream):
        b.demod_device(buf, "cs16", B, d_out, soft=True, in_stride=B, out_stride=maxf * b.Nbits, stream=stream)

    def one_step():
        bs.step(cap, T, consume)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    for _ in range(max(args.warmup, 3)):            # warms up the NCCL communicator too
        one_step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = b.launches
    e0.record(comp_s)
    for _ in range(args.steps):
        one_step()
    e1.record(comp_s)
    barrier()
    ms = e0.elapsed_time(e1)
    frames, consumed = b.counts()
    tv = torch.tensor([ms], dtype=torch.float64, device=dev)
    sv = torch.tensor([float(consumed.sum())], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tv, op=dist.ReduceOp.MAX); dist.all_reduce(sv, op=dist.ReduceOp.SUM)
    if rank == 0:
        fed = float(world) * C * T * B * args.steps                       # samples scattered and demodulated
        value = fed / (float(tv[0]) * 1e-3) / 1e6
        sent = float(world - 1) * C * T * B * 4 * args.steps              # bytes that left rank 0 over NVLink
        line = {"metric": METRIC, "mode": "scatter", "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": float(tv[0]) / args.steps, "higher_is_better": True, "scaling": "weak", "dtype": "f32", "data": "synthetic",
                "config": {"workload": f"NCCL scatter of a cs16 capture held by rank 0: {world} x {C} RS41-prod channels, {T} time blocks of {B} samples per step, "
                                       "scatter of block t+1 overlapped with the demodulation of block t",
                           "consumed_last_block_samples": float(sv[0])},
                "collective": {"op": "torch.distributed.scatter (NCCL grouped send/recv), src rank 0", "bytes_out_of_rank0_per_step": sent / args.steps,
                               "rate_GBs": sent / (float(tv[0]) * 1e-3) / 1e9, "nvlink_peer_copy_GBs_measured": 770.0,
                               "note": "rate = bytes leaving rank 0 / whole step time (the scatter overlaps the kernels; it is not timed alone)"},
                "gpu_launches": int(b.launches - l0), "host": numa}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


# ---------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------
def run_ours(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    nchan, nsamp = args.channels, args.nsamp
    iq = gen_channels(rank, nchan, nsamp)                                   # int16 [nchan, nsamp, 2]
    parity_chans = sorted(set(list(range(min(8, nchan))) + list(range(max(0, nchan - 8), nchan)) + ([NOISE_CHANNEL] if nchan > NOISE_CHANNEL else [])))

    # CPU baseline first (forked worker processes: before this process touches CUDA)
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        r = time_reference(iq, repeats=3)
        kind, what = baseline_kind()
        cpu = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": kind, "cpu_model": cpu_model(), "passes": r["passes"],
               "logical_cpus": r["logical_cpus"], "cpu_quota": r["cpu_quota"],
               "per_core_median": r["per_core_median"], "one_core": one_core_reference(iq),
               "sample": f"the whole workload ({nchan} ch x {nsamp} samples) x 3 passes, median; one pinned process per usable CPU "
                         f"(min of the affinity mask and the container's CFS quota); {what}"}
    ref_first = reference_outputs(iq, parity_chans)                          # a stream that starts at sample 0 (first e2e push)
    ref_last = reference_outputs(iq, parity_chans, args.warmup + args.steps)   # the handle's (warmup + steps)-th walk over the buffer

    import torch
    import torch.distributed as dist
    from radiosonde_auto_rx_b200 import fsk
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the demodulator has no CPU fallback)")
    torch.cuda.set_device(local)
    numa = bind_to_gpu_numa(local)                   # before the pinned host buffers of the end-to-end leg are allocated
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    # HBM-resident float2 input holding exactly (float)int16/1000 (fsk_demod.c:305-306)
    d_in = torch.empty((nchan, nsamp, 2), dtype=torch.float32, device="cuda")
    for c0 in range(0, nchan, 32):                                          # numpy: IEEE division, same as the CLI's
        d_in[c0:c0 + 32] = torch.from_numpy(iq[c0:c0 + 32].astype(np.float32) / np.float32(1000)).cuda()
    b = fsk.FskBatch(FS, RS, M, P, NSYM, 1000, MASK, LO, HI, 1, nchan=nchan, device=local)
    maxf = b.max_frames(nsamp)
    d_out = torch.zeros((nchan, maxf * b.Nbits), dtype=torch.float32, device="cuda")
    stream = torch.cuda.Stream()          # the launching stream: events below are recorded on it (never the NULL stream)
    torch.cuda.synchronize()

    def step():
        b.demod_device(d_in, "f32", nsamp, d_out, soft=True, in_stride=nsamp, out_stride=maxf * b.Nbits, stream=stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    clocks = Clocks(local) if rank == 0 else None
    for _ in range(args.warmup):
        step()
    barrier()
    if rank == 0:
        time.sleep(0.3)                    # let the sampler reach steady state before the timed region
    barrier()
    launches0 = b.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.time()
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    barrier()
    t_wall1 = time.time()
    ms = e0.elapsed_time(e1)
    clk = clocks.stop(t_wall0, t_wall1) if clocks else None
    frames, consumed = b.counts()
    samples_per_step = float(consumed.sum())
    launches = b.launches - launches0
    kernel_name = b.last_kernel()
    # parity gate, device-resident leg: the output of the LAST timed step
    out_sel = {c: d_out[c].cpu().numpy() for c in parity_chans}
    bad_dev = count_mismatches(ref_last, lambda c: out_sel[c], frames, b.Nbits)

    # ---- end to end through the host-buffer C-ABI (pinned cs16 in, soft bits out) ----
    h_in = torch.from_numpy(iq).pin_memory()
    be = fsk.FskBatch(FS, RS, M, P, NSYM, 1000, MASK, LO, HI, 1, nchan=nchan, device=local)
    rows = int(be.L.b200fsk_push_frames_bound(be.h, nsamp + be.nin_max, 0)) * be.Nbits
    h_out = torch.zeros((nchan, rows), dtype=torch.float32).pin_memory()
    h_in_np, h_out_np = h_in.numpy(), h_out.numpy()
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    # every push continues the stream of the previous one; the parity gate needs a stream that starts at sample 0: check the FIRST push
    _, fr0 = be.push(h_in_np.reshape(nchan, -1), "cs16", soft=True, out=h_out_np)
    bad_e2e = count_mismatches(ref_first, lambda c: h_out_np[c], fr0, be.Nbits)
    for _ in range(max(0, min(2, args.warmup) - 1)):
        be.push(h_in_np.reshape(nchan, -1), "cs16", soft=True, out=h_out_np)
    barrier()
    t0 = time.perf_counter()
    e2e_samples = 0.0
    for _ in range(e2e_steps):
        _, fr = be.push(h_in_np.reshape(nchan, -1), "cs16", soft=True, out=h_out_np)
        e2e_samples += float(nsamp) * nchan
    torch.cuda.synchronize()
    e2e_dt = time.perf_counter() - t0
    d2h = float(fr.max()) * be.Nbits * 4 * nchan
    e2e_launches = be.launches

    # ---- channel-count sweep (rank 0, single GPU): device-resident, channels beyond the generated batch repeat it ----
    sweep = None
    if world == 1 and not args.no_sweep:
        sweep = [{"channels": nchan, "nsamp": nsamp, "kernel": kernel_name, "value": samples_per_step * args.steps / (ms * 1e-3) / 1e6,
                  "parity_mismatches": bad_dev, "parity_channels": len(parity_chans)}]
        ns = min(nsamp, args.sweep_nsamp)
        sub = sorted(set([0, 1, 2, NOISE_CHANNEL, nchan - 2, nchan - 1])) if nchan > NOISE_CHANNEL else list(range(nchan))
        ref_s = reference_outputs(iq[:, :ns], sub, 3 + args.sweep_steps)     # 3 warm-up launches + the timed ones on one handle
        del d_out
        for nc in args.sweep_channels:
            if nc == nchan:
                continue
            reps = (nc + nchan - 1) // nchan
            x = d_in[:, :ns].repeat(reps, 1, 1)[:nc].contiguous()
            bs = fsk.FskBatch(FS, RS, M, P, NSYM, 1000, MASK, LO, HI, 1, nchan=nc, device=local)
            mf = bs.max_frames(ns)
            y = torch.zeros((nc, mf * bs.Nbits), dtype=torch.float32, device="cuda")
            for _ in range(3):
                bs.demod_device(x, "f32", ns, y, soft=True, in_stride=ns, out_stride=mf * bs.Nbits, stream=stream)
            torch.cuda.synchronize()
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s0.record(stream)
            for _ in range(args.sweep_steps):
                bs.demod_device(x, "f32", ns, y, soft=True, in_stride=ns, out_stride=mf * bs.Nbits, stream=stream)
            s1.record(stream)
            torch.cuda.synchronize()
            fr_s, cons_s = bs.counts()
            # parity: the last replica of the checked channels (highest channel indices of the big batch) against the reference
            base = (nc - 1) // nchan * nchan
            chk = {c: ref_s[c] for c in sub if base + c < nc}
            ysel = {c: y[base + c].cpu().numpy() for c in chk}
            bad = count_mismatches(chk, lambda c: ysel[c], {c: fr_s[base + c] for c in chk}, bs.Nbits)
            sweep.append({"channels": nc, "nsamp": ns, "kernel": bs.last_kernel(), "value": float(cons_s.sum()) * args.sweep_steps / (s0.elapsed_time(s1) * 1e-3) / 1e6,
                          "parity_mismatches": bad, "parity_channels": len(chk)})
            bs.close()
            del x, y
            torch.cuda.empty_cache()
        sweep.sort(key=lambda d: d["channels"])

    # ---- max over ranks ----
    tv = torch.tensor([ms, e2e_dt], dtype=torch.float64, device="cuda")
    sv = torch.tensor([samples_per_step, e2e_samples, float(bad_dev), float(bad_e2e)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tv, op=dist.ReduceOp.MAX)
        dist.all_reduce(sv, op=dist.ReduceOp.SUM)
    ms_max, e2e_dt_max = float(tv[0]), float(tv[1])
    tot_samples, tot_e2e = float(sv[0]), float(sv[1])
    bad_dev_all, bad_e2e_all = int(sv[2]), int(sv[3])
    sweep_bad = sum(s["parity_mismatches"] for s in sweep) if sweep else 0

    rc = 0
    if rank == 0:
        parity = {"channels": len(parity_chans) * world, "mismatches": bad_dev_all + bad_e2e_all, "device_resident_mismatches": bad_dev_all,
                  "e2e_mismatches": bad_e2e_all, "checked_against": baseline_kind()[1],
                  "what": f"soft bits of channels {parity_chans} of every rank, all frames of the 2^{int(np.log2(nsamp))}-sample stream, bit for bit: "
                          "the last timed device-resident step (the reference modem object walks the buffer warmup+steps times, state carried over, "
                          "exactly like the handle) and the first end-to-end push"}
        if bad_dev_all or bad_e2e_all or sweep_bad:
            print(json.dumps({"metric": METRIC, "error": "parity gate failed: outputs differ from the reference modem; no value reported",
                              "parity": parity, "sweep": sweep}), flush=True)
            rc = 1
        else:
            value = tot_samples * args.steps / (ms_max * 1e-3) / 1e6
            e2e_val = tot_e2e / e2e_dt_max / 1e6
            peak, peak_src = measured_peak()
            kern_ms = ms / args.steps                       # one launch per step on this rank
            achieved = samples_per_step * BYTES_PER_SAMPLE_F32 / (kern_ms * 1e-3) / 1e9
            traffic, traffic_src = committed_traffic(kernel_name, nchan, nsamp)
            line = {
                "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": workload_name(nchan, nsamp), "channels_total": nchan * world,
                           "concurrent_48k_channels_realtime": value * 1e6 / FS,
                           "input": "float2 [channel][sample] resident in HBM, 2 GiB/GPU > L2 (no flush needed)",
                           "parallelism": f"channel-sharded x{world}, no data-path collective",
                           "sweep": sweep,
                           "sweep_note": "Msamples/s vs channels per GPU, device-resident; the library picks the frame kernel by channels per SM: up to 4 the "
                                         "duo kernel (one oscillator warp that owns a sub-partition + specialised warps per channel, two channels per CTA; "
                                         "bound by the oscillator recurrence + role coupling), beyond that 15 channels per CTA sharing one lane-per-chain "
                                         "oscillator warp (fsk_wpc_kernel, bound by instruction issue)"},
                "clocks": clk,
                "parity": parity,
                "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": float(nsamp) * nchan * 4,
                        "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                        "api": "b200fsk_push_host: pinned cs16 host buffer -> H2D -> kernel -> D2H soft bits (synchronous)", "host": numa},
                "gpu_launches": int(launches),
                "gpu_launches_e2e": int(e2e_launches),
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                             "traffic": traffic, "traffic_source": traffic_src, "kernel": kernel_name, "kernel_ms": kern_ms,
                             "algorithmic_bytes_per_sample": BYTES_PER_SAMPLE_F32, "peak_source": peak_src,
                             "note": "not HBM-bound: ~125 un-fusable FP32 operations per sample (no FMA allowed: bit-exactness) put the issue ceiling near 0.3 Tsps; "
                                     "at 256 channels (2 per SM on 128 of 148 SMs) the per-channel oscillator recurrence (one sample per >= 10.7 SM cycles, "
                                     "13.6 measured inside the kernel) caps the batch at ~47 Gsps; see DESIGN.md 4.1 and profiles/r02_notes.md"},
            }
            if sweep:
                best = max(sweep, key=lambda d: d["value"])
                line["roofline"]["best_of_sweep"] = {"channels": best["channels"], "kernel": best["kernel"], "value": best["value"],
                                                     "achieved": best["value"] * 1e6 * BYTES_PER_SAMPLE_F32 / 1e9,
                                                     "frac": best["value"] * 1e6 * BYTES_PER_SAMPLE_F32 / 1e9 / peak}
            if cpu:
                line["cpu_baseline"] = cpu
            if world == 1 and not args.no_secondary:
                try:
                    del d_in
                    torch.cuda.empty_cache()
                    line["secondary"] = run_dft_secondary(args, local)
                except Exception as e:                        # noqa: BLE001  -- the headline line must not die with the secondary workload
                    line["secondary"] = {"error": f"{type(e).__name__}: {e}"}
            print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return rc


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--channels", type=int, default=256)
    ap.add_argument("--nsamp", type=int, default=1 << 20)
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sweep", action="store_true")
    ap.add_argument("--sweep-channels", type=int, nargs="*", default=[256, 1024, 4096])
    ap.add_argument("--sweep-nsamp", type=int, default=1 << 18)
    ap.add_argument("--sweep-steps", type=int, default=3)
    ap.add_argument("--scatter", action="store_true", help="time the NCCL scatter + demod path (rank 0 holds the capture) instead of the resident-input bench")
    ap.add_argument("--scatter-channels", type=int, default=128, help="channels per GPU (BASELINE configs[3]: 1024 channels over 8 GPUs)")
    ap.add_argument("--scatter-block", type=int, default=1 << 17)
    ap.add_argument("--scatter-blocks", type=int, default=8)
    ap.add_argument("--no-secondary", action="store_true", help="skip the dft_detect workload (BASELINE configs[2])")
    ap.add_argument("--dft-channels", type=int, default=4096)
    ap.add_argument("--dft-nsamp", type=int, default=1 << 18)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    if args.impl == "reference":
        run_reference(args)
        return 0
    if args.warmup < 3:
        args.warmup = 3
    if args.scatter:
        return run_scatter(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
