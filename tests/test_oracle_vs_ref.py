"""CPU tests that PIN the oracle: our C restatements vs the UNMODIFIED reference compiled into oracle/_ref
(fsk: bit-identical outputs and state; dft_detect: type/sign/position exact, score 1e-3).  Needs oracle/_ref,
i.e. /root/reference at build time or the prebuilt files that travel with the tree."""
import numpy as np
import pytest

from oracle import pyoracle as po
from radiosonde_auto_rx_b200 import siggen

needs_ref = pytest.mark.skipif(not po.ref_available(), reason="oracle/_ref not built and /root/reference absent")

FSK_CASES = {
    "rs41": (po.RS41, dict()),
    "rs41_noise": (po.RS41, None),
    "default_hard": (po.FskConfig(48000, 4800, 2, 10, 50, soft=False), dict(ppm=10.0)),
    "rs92": (po.FskConfig(48000, 4800, 2, 10, 50, -20000, 20000), dict()),
    "dfm_softinv": (po.FskConfig(50000, 2500, 2, 10, 50, -5000, 5000, softinv=True), dict(Fs=50000, Rs=2500, spacing=5000.0)),
    "m10": (po.FskConfig(48080, 9616, 2, 5, 50, -10000, 10000), dict(Fs=48080, Rs=9616)),
    "wxr301_mask": (po.FskConfig(96000, 4800, 2, 10, 50, -40000, 40000, 50000), dict(Fs=96000, Rs=4800, spacing=50000.0)),
    "4fsk_mask": (po.FskConfig(48000, 4800, 4, 10, 50, -20000, 20000, mask=4800), dict(M=4, spacing=4800.0)),
    "4fsk_peak_hard": (po.FskConfig(48000, 4800, 4, 10, 50, soft=False), dict(M=4, spacing=4800.0)),
}


@needs_ref
@pytest.mark.parametrize("name", list(FSK_CASES))
def test_fsk_oracle_bit_identical_to_reference(built, name):
    cfg, kw = FSK_CASES[name]
    n = 200_000
    iq = siggen.noise_iq(77, n) if kw is None else siggen.fsk_iq(40 + len(name), n, ebno_db=9.0, **kw)
    out, cons, recs = po.Oracle(cfg).run_stream(iq)
    rout, rcons, rrec, rSf = po.ref_run_stream(cfg, iq)
    assert cons == rcons and len(recs) == len(rrec) > 0
    assert out.tobytes() == rout.tobytes()
    assert np.array_equal(recs["nin_next"], rrec[:, 1].astype(np.int32))
    for k, col in (("norm_rx_timing", 2), ("ppm", 3), ("EbNodB", 4), ("snr_est", 5)):
        assert np.array_equal(recs[k].view(np.uint32), rrec[:, col].view(np.uint32)), k
    assert np.array_equal(recs["f_est"].view(np.uint32), rrec[:, 6:10].view(np.uint32))


@needs_ref
def test_fsk_oracle_matches_reference_cli_bytes(built):
    """The actual reference executable on stdin/stdout (the drop-in surface) vs the oracle's frame pump."""
    cfg = po.RS41
    iq = siggen.fsk_iq(5, 100_000, ebno_db=12.0, ppm=25.0)
    out, _, _ = po.Oracle(cfg).run_stream(iq)
    so, se, rc = po.ref_cli_fsk_demod(cfg, iq.tobytes())
    assert rc == 0 and so == out.tobytes()
    assert se.startswith(b"Setting estimator limits to -5000 to 5000 Hz.")


@needs_ref
@pytest.mark.parametrize("kind", list(siggen.HEADER_BANK) + ["noise"])
def test_dft_oracle_matches_reference_binary(built, kind):
    n = 100_000
    iq = siggen.noise_iq(9, n) if kind == "noise" else siggen.fm_header_iq(200 + len(kind), n, kind, snr_db=16.0)
    rc, dets, _ = po.dft_oracle_run(iq, 48000, 15.0, True, True)
    rrc, rdets = po.ref_cli_dft_detect(iq, 48000, 15.0, True, True)
    assert rc == rrc and len(dets) == len(rdets)
    if kind != "noise":
        assert len(dets) > 0 and dets[0][0].startswith(kind[:3]) or kind == "M10"
    for a, b in zip(dets, rdets):
        assert a[0] == b[0] and a[2] == b[2] and abs(a[1] - b[1]) < 1e-3 and abs(a[3] - b[3]) < 0.2


@needs_ref
@pytest.mark.parametrize("sr,bw,hz,dtype", [(2400000, 15.0, 300e3, np.int16), (2400000, 15.0, -500e3, np.uint8), (1000000, 15.0, 200e3, np.float32),
                                            (960000, 100.0, 100e3, np.int16), (1000000, 15.0, 0.0, np.int16), (48000, 15.0, 0.0, np.int16)])
def test_dft_wideband_oracle_matches_reference_binary(built, sr, bw, hz, dtype):
    """`dft_detect --IQ fq`: IF-rate walk-up (1 MHz -> 50 kHz), wide IF (--bw 100 -> 120 kHz, the 60 kHz filter branch), no decimation
    (sr = 48 kHz: decM 1, 7-tap filter), all three sample formats"""
    emit = [("RS41", 300e3 if sr > 1e6 else (200e3 if sr == 1000000 else (100e3 if sr == 960000 else 0.0))), ("DFM9", -500e3 if sr > 1e6 else -300e3)]
    if sr == 48000:
        emit = [("RS41", 0.0)]
    wide = siggen.wideband_mix(41, int(sr * 1.3), sr, emit, snr_db=8.0, dc=(120.0, -60.0))
    if dtype == np.float32:
        wide = wide.astype(np.float32) / np.float32(32768.0)
    elif dtype == np.uint8:
        wide = np.clip(np.rint(wide.astype(np.float32) / 96.0) + 128, 0, 255).astype(np.uint8)
    bits = {np.int16: 16, np.uint8: 8, np.float32: 32}[dtype]
    rrc, rdets = po.ref_cli_dft_detect(wide, sr, bw, True, True, bits=bits, fq=hz / sr)
    rc, dets, info, _ = po.dft_oracle_run_wideband(wide, sr, hz / sr, bw, True, True)
    assert rc == rrc and len(dets) == len(rdets), (info, rc, rrc, dets, rdets)
    for a, b in zip(dets, rdets):
        assert a[0] == b[0] and a[2] == b[2] and abs(a[1] - b[1]) < 1e-3 and abs(a[3] - b[3]) < 0.2
    if (sr, bw) == (960000, 100.0):
        assert info["if_sr"] == 120000 and info["decM"] == 8
    if sr == 48000:
        assert info["decM"] == 1 and len(dets) > 0


def _random_fsk_config(seed):
    """a valid modem configuration drawn the way fsk_create_hbr accepts it (utils/fsk.c:120-130: Fs % Rs == 0, (Fs/Rs) % P == 0,
    P >= 4..., M in {2,4}) with matching signal parameters"""
    rng = np.random.default_rng(seed)
    Rs = int(rng.choice([1200, 2400, 2500, 4800, 9600]))
    Ts = int(rng.choice([8, 10, 16, 20, 40]))
    Fs = Rs * Ts
    P = int(rng.choice([p for p in (4, 5, 8, 10, 16, 20) if Ts % p == 0 and p <= Ts]))
    M = int(rng.choice([2, 4]))
    Nsym = int(rng.choice([32, 50, 64, 100, 300]))
    spacing = float(Rs * rng.choice([1.0, 1.5, 2.0]))
    span = spacing * (M - 1)
    lim = int(min(Fs // 2 - 1, span + 2 * Rs))
    mask = int(spacing) if rng.random() < 0.5 else 0
    cfg = po.FskConfig(Fs, Rs, M, P, Nsym, -lim, lim, mask, soft=bool(rng.random() < 0.7), softinv=bool(rng.random() < 0.3))
    sig = dict(Fs=Fs, Rs=Rs, M=M, spacing=spacing, ebno_db=float(rng.uniform(6, 16)), ppm=float(rng.uniform(-40, 40)))
    return cfg, sig


@needs_ref
@pytest.mark.parametrize("seed", range(100, 124))
def test_fsk_oracle_bit_identical_on_random_configurations(built, seed):
    """24 seeded modem configurations beyond the production lines (rates, oversampling, P, M, Nsym, estimator, soft/hard/inverted)"""
    cfg, sig = _random_fsk_config(seed)
    n = 40 * (cfg.Nsym + 1) * (cfg.Fs // cfg.Rs)
    iq = siggen.fsk_iq(seed, n, **sig)
    out, cons, recs = po.Oracle(cfg).run_stream(iq)
    rout, rcons, rrec, _ = po.ref_run_stream(cfg, iq)
    assert cons == rcons and len(recs) == len(rrec) > 10, (cfg, cons, rcons)
    assert out.tobytes() == rout.tobytes(), cfg
    assert np.array_equal(recs["nin_next"], rrec[:, 1].astype(np.int32))
    assert np.array_equal(recs["f_est"].view(np.uint32), rrec[:, 6:10].view(np.uint32))
    assert np.array_equal(recs["snr_est"].view(np.uint32), rrec[:, 5].view(np.uint32))
