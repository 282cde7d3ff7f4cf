#!/usr/bin/env python
"""Generate the golden fixtures in this directory by running the UNMODIFIED reference (oracle/_ref, compiled from
/root/reference by oracle/Makefile) on small seeded inputs.  Run in the build container:

    python tests/golden/make_golden.py

Outputs (committed): golden_fsk.npz, golden_dft.json, golden_dft_wb.json, golden_rs41.json.  The inputs are regenerated from seeds by
radiosonde_auto_rx_b200.siggen, so only outputs + an input checksum are stored.
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyoracle as po                      # noqa: E402
from radiosonde_auto_rx_b200 import siggen             # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

FSK_CASES = {
    # name: (FskConfig kwargs, siggen kwargs, nsamp)
    "rs41_prod": (dict(Fs=48000, Rs=4800, M=2, P=5, Nsym=300, lower=-5000, upper=5000, mask=5000), dict(seed=11, ebno_db=10.0, ppm=35.0), 60000),
    "rs41_noise": (dict(Fs=48000, Rs=4800, M=2, P=5, Nsym=300, lower=-5000, upper=5000, mask=5000), None, 40000),
    "default_hard": (dict(Fs=48000, Rs=4800, M=2, P=10, Nsym=50, soft=False), dict(seed=12, ebno_db=9.0, ppm=-20.0), 30000),
    "dfm_softinv": (dict(Fs=50000, Rs=2500, M=2, P=10, Nsym=50, lower=-5000, upper=5000, softinv=True), dict(seed=13, Fs=50000, Rs=2500, spacing=5000.0, ebno_db=12.0), 30000),
    "4fsk_mask": (dict(Fs=48000, Rs=4800, M=4, P=10, Nsym=50, lower=-20000, upper=20000, mask=4800), dict(seed=14, M=4, spacing=4800.0, ebno_db=13.0), 30000),
}
DFT_CASES = {k: dict(seed=700 + i, kind=k, nsamp=60000, snr_db=16.0) for i, k in enumerate(["RS41", "DFM9", "M10", "MTS01", "WXR301", "noise"])}


# 96 kHz IF (auto_rx wideband_sondes: `dft_detect ... --iq --bw 64 --dc - 96000 16`, scan.py:476-482; everything in the detector doubles:
# N_DFT 16384, K 11264, FIR taps 97/193): name -> (kind, --bw kHz, FM deviation Hz, seed)
DFT96_CASES = {"WXR301_bw64": ("WXR301", 64.0, 12000.0, 801), "WXRPN9_bw64": ("WXRPN9", 64.0, 12000.0, 802), "RS41_bw32": ("RS41", 32.0, 2400.0, 803),
               "DFM9_bw64": ("DFM9", 64.0, 2400.0, 804), "noise_bw64": ("noise", 64.0, 0.0, 805)}


def dft96_input(name):
    kind, _, dev, seed = DFT96_CASES[name]
    return siggen.noise_iq(seed, 240000) if kind == "noise" else siggen.fm_header_iq(seed, 240000, kind, Fs=96000, dev_hz=dev, snr_db=17.0)


def fsk_input(name):
    ck, sk, n = FSK_CASES[name]
    if sk is None:
        return siggen.noise_iq(99, n)
    sk = dict(sk)
    return siggen.fsk_iq(sk.pop("seed"), n, **sk)


def dft_input(name):
    c = DFT_CASES[name]
    return siggen.noise_iq(c["seed"], c["nsamp"]) if c["kind"] == "noise" else siggen.fm_header_iq(c["seed"], c["nsamp"], c["kind"], snr_db=c["snr_db"])


# wideband captures for the `--IQ <fq>` mode: name -> (sr_base, seconds, seed, emitters, candidate Hz list)
_M10_TAIL = "".join("10" if b == "1" else "01" for b in "0110010010011111")
DFT_WB_CASES = {
    "wb2400k": (2400000, 1.0, 11, [("RS41", 300e3), ("DFM9", -500e3), ("M10", 800e3, _M10_TAIL), ("RS92", -900e3)],
                [300e3, -500e3, 800e3, -900e3, 100e3]),
    "wb1000k": (1000000, 1.6, 12, [("RS41", 200e3), ("LMS6", -300e3), ("IMET5", 50e3)], [200e3, -300e3, 50e3, -123456.0]),   # IF 50 kHz
}


# soft-bit streams of tests/rs41_frames.py: (seed, noise sigma, burst) -- the cases of tests/test_ecc.py
RS41_CASES = [(1, 0.30, None), (2, 0.40, None), (3, 0.45, None), (4, 0.50, None), (5, 0.38, (2, 0x39, 2)), (6, 0.36, (1, 0x65, 30)),
              (7, 0.2, (3, 200, 40)), (8, 0.42, None), (9, 0.43, None), (10, 0.41, None)]


def dft_wb_input(name):
    sr, secs, seed, emit, _ = DFT_WB_CASES[name]
    return siggen.wideband_mix(seed, int(sr * secs), sr, emit, snr_db=6.0, dc=(250.0, -150.0))


def main():
    out = {}
    for name, (ck, _, _) in FSK_CASES.items():
        cfg = po.FskConfig(**ck)
        iq = fsk_input(name)
        so, se, rc = po.ref_cli_fsk_demod(cfg, iq.tobytes())          # the reference EXECUTABLE, stdin -> stdout
        assert rc == 0
        o, cons, recs, Sf = po.ref_run_stream(cfg, iq)                  # same modem through its C API, plus state
        assert o.tobytes() == so
        out[name + "/out"] = np.frombuffer(so, dtype=np.float32 if cfg.soft else np.uint8)
        out[name + "/nin"] = recs[:, 1].astype(np.int32)
        out[name + "/ppm"] = recs[:, 3].copy()
        out[name + "/snr_est"] = recs[:, 5].copy()
        out[name + "/f_est"] = recs[:, 6:10].copy()
        out[name + "/sha"] = np.frombuffer(hashlib.sha256(iq.tobytes()).digest(), dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, "golden_fsk.npz"), **out)
    dj = {}
    for name in DFT_CASES:
        iq = dft_input(name)
        rc, dets = po.ref_cli_dft_detect(iq, 48000, 15.0, True, True)
        dj[name] = {"exit_code": rc, "dets": dets, "sha": hashlib.sha256(iq.tobytes()).hexdigest()}
    json.dump(dj, open(os.path.join(HERE, "golden_dft.json"), "w"), indent=1)
    d96 = {}
    for name, (kind, bw, _, _) in DFT96_CASES.items():
        iq = dft96_input(name)
        rc, dets = po.ref_cli_dft_detect(iq, 96000, bw, True, True)      # dft_detect -v -c --iq --bw <bw> --dc - 96000 16
        d96[name] = {"exit_code": rc, "dets": dets, "bw": bw, "sha": hashlib.sha256(iq.tobytes()).hexdigest()}
    json.dump(d96, open(os.path.join(HERE, "golden_dft96.json"), "w"), indent=1)
    wj = {}
    for name, (sr, _, _, _, hz) in DFT_WB_CASES.items():
        iq = dft_wb_input(name)
        wj[name] = {"sha": hashlib.sha256(iq.tobytes()).hexdigest(), "fq": {}}
        for f in hz:
            rc, dets = po.ref_cli_dft_detect(iq, sr, 15.0, True, True, fq=f / sr)     # dft_detect -v -c --IQ fq --bw 15 --dc - sr 16
            wj[name]["fq"][repr(f)] = {"exit_code": rc, "dets": dets}
    json.dump(wj, open(os.path.join(HERE, "golden_dft_wb.json"), "w"), indent=1)
    # RS41 frame sync + Reed-Solomon stage: what `rs41mod -r [--ecc2] --softin -i` prints for seeded soft-bit streams
    sys.path.insert(0, os.path.dirname(HERE))
    import rs41_frames as rf
    rj = {}
    for seed, sigma, burst in RS41_CASES:
        soft, _ = rf.soft_stream(seed, 8, sigma, burst=burst)
        ent = {"sha": hashlib.sha256(soft.tobytes()).hexdigest()}
        for ecc in (1, 2):
            ref = po.ref_cli_rs41mod_raw(soft.tobytes(), inv=True, ecc=ecc)
            ent[f"ecc{ecc}"] = [[hashlib.sha1(h.encode()).hexdigest()[:16], t.strip()] for h, t in ref]
        rj[str(seed)] = ent
    json.dump(rj, open(os.path.join(HERE, "golden_rs41.json"), "w"), indent=1)
    print("wrote golden_fsk.npz, golden_dft.json, golden_dft96.json, golden_dft_wb.json, golden_rs41.json")


if __name__ == "__main__":
    main()
