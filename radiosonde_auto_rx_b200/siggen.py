"""Seeded synthetic IQ generators for the IQ->bits hot path (tests + bench).

The reference ships no recorded IQ (fixtures are external downloads), so every
workload is synthesised here, deterministically from ``numpy.random.default_rng(seed)``, and
QUANTISED TO cs16 -- the byte format the reference CLIs read on stdin (utils/fsk_demod.c:281-309,
scan/dft_detect.c:556-591) -- so the oracle and the CUDA path see bit-identical input.

Nothing here is on the product path; it only manufactures input.
"""
from __future__ import annotations

import numpy as np

# 64-bit RS41 frame header as the reference detector spells it (scan/dft_detect.c:46-47).
RS41_HEADER = "0000100001101101010100111000100001000100011010010100100000011111"


def _bits_with_header(rng: np.random.Generator, nbits: int, header: str | None, period: int) -> np.ndarray:
    bits = rng.integers(0, 2, size=nbits, dtype=np.uint8)
    if header:
        h = np.frombuffer(header.encode(), dtype=np.uint8) - ord("0")
        for s in range(int(rng.integers(0, period)), nbits - len(h), period):
            bits[s:s + len(h)] = h
    return bits


def fsk_iq(seed: int, nsamp: int, Fs: int = 48000, Rs: int = 4800, M: int = 2,
           f_lo: float | None = None, spacing: float | None = None, ebno_db: float = 15.0,
           ppm: float = 0.0, amp: float = 8000.0, header: str | None = RS41_HEADER,
           header_period: int = 4800, return_bits: bool = False, bits: np.ndarray | None = None):
    """Continuous-phase M-FSK at complex baseband -> cs16 array of shape (nsamp, 2), dtype int16.

    Tone m sits at ``f_lo + m*spacing`` Hz.  ``ppm`` stretches the symbol clock, ``ebno_db`` sets
    AWGN (complex, total variance N0 per sample; Es = amp^2 * Ts).  All randomness comes from ``seed``.
    """
    rng = np.random.default_rng(seed)
    if spacing is None:
        spacing = float(Rs)
    if f_lo is None:
        f_lo = -spacing * (M - 1) / 2.0 + float(rng.uniform(-2000.0, 2000.0))
    bps = 1 if M == 2 else 2
    Ts = Fs / Rs
    nsym = int(nsamp / Ts * (1 + abs(ppm) * 1e-6)) + 4
    if bits is None:
        bits = _bits_with_header(rng, nsym * bps, header, header_period)
    else:                                            # caller-supplied payload (e.g. well-formed RS41 frames); padded with random bits
        given = np.asarray(bits, dtype=np.uint8)
        bits = rng.integers(0, 2, size=nsym * bps, dtype=np.uint8)
        bits[: min(given.size, bits.size)] = given[: bits.size]
    if M == 2:
        sym = bits.astype(np.int64)
    else:
        sym = (bits[0::2].astype(np.int64) << 1) | bits[1::2]
    tau0 = float(rng.uniform(0.0, 1.0))
    sidx = (np.arange(nsamp, dtype=np.float64) * ((1.0 + ppm * 1e-6) / Ts) + tau0).astype(np.int64)
    f = (f_lo + spacing * sym)[sidx]
    phase = np.cumsum(f)
    phase *= 2.0 * np.pi / Fs
    phase += float(rng.uniform(0, 2 * np.pi))
    np.remainder(phase, 2.0 * np.pi, out=phase)
    ph32 = phase.astype(np.float32)
    ebno = 10.0 ** (ebno_db / 10.0)
    n0 = amp * amp * Ts / (ebno * bps)
    sigma = np.float32(np.sqrt(n0 / 2.0))
    out = rng.standard_normal((nsamp, 2), dtype=np.float32)
    out *= sigma
    out[:, 0] += np.float32(amp) * np.cos(ph32)
    out[:, 1] += np.float32(amp) * np.sin(ph32)
    np.rint(out, out=out)
    np.clip(out, -32768, 32767, out=out)
    iq = out.astype(np.int16)
    if return_bits:
        return iq, bits
    return iq


def noise_iq(seed: int, nsamp: int, sigma: float = 3000.0) -> np.ndarray:
    """Noise-only channel (cs16)."""
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((nsamp, 2), dtype=np.float32) * sigma
    return np.clip(np.rint(x), -32768, 32767).astype(np.int16)


def fsk_batch(base_seed: int, nchan: int, nsamp: int, workers: int = 8, **kw) -> np.ndarray:
    """``nchan`` independent channels -> int16 [nchan, nsamp, 2]; channel c uses seed base_seed+c,
    a per-channel clock offset in +-50 ppm and Eb/N0 cycling over {10, 14, 20, 40} dB unless given."""
    from concurrent.futures import ThreadPoolExecutor

    out = np.empty((nchan, nsamp, 2), dtype=np.int16)

    def one(c: int) -> None:
        k = dict(kw)
        r = np.random.default_rng(base_seed * 7919 + c)
        k.setdefault("ppm", float(r.uniform(-50.0, 50.0)))
        k.setdefault("ebno_db", (10.0, 14.0, 20.0, 40.0)[c % 4])
        out[c] = fsk_iq(base_seed + c, nsamp, **k)

    with ThreadPoolExecutor(max_workers=workers) as ex:
        list(ex.map(one, range(nchan)))
    return out


# ---------------------------------------------------------------------------------------------------
# dft_detect workloads: FM (GFSK) carriers whose bit stream contains the sonde-type header
# ---------------------------------------------------------------------------------------------------
# (sps, header) of the reference's header bank, scan/dft_detect.c:43-126,177-194 (index = rs_hdr[] row)
HEADER_BANK = {
    "DFM9": (0, 2500, "10011010100110010101101001010101"),
    "RS41": (1, 4800, RS41_HEADER),
    "RS92": (2, 4800, "10100110011001101001" * 2 + "1010011001100110100110101010100110101001"),
    "LMS6": (3, 4800, "0101011000001000" "0001110010010111" "0001101010100111" "0011110100111110"),
    "IMET5": (4, 4800, "0000000001" "0101010101" "0001001001" "0001001001"),
    "MK2LMS": (5, 9616, "0010100111" "0010100111" "0001001001" "0010010101"),
    "M10": (6, 9608, "1001100110010100110010011001" "1010"),
    "MEISEI": (7, 2400, "110011001101001101001101010100101010110010101010"),
    "RD94RD41": (8, 4800, "10100110010110101001" "10010101011010010101" "10101001010101010101" "10011001010110101001"),
    "MRZ": (9, 2400, "1001100110011001" "1001101010101010"),
    "MTS01": (10, 1200, "10101010" "10101010" "10110100" "00101011"),
    "WXR301": (12, 4800, "10101010" * 3 + "00101101" "11010100"),
    "WXRPN9": (13, 5000, "10101010" * 3 + "11000001" "10010100"),
}


def fm_header_iq(seed: int, nsamp: int, kind: str, Fs: int = 48000, dev_hz: float = 2400.0, snr_db: float = 20.0,
                 f_off: float | None = None, invert: bool = False, period_bits: int | None = None, amp: float = 8000.0,
                 tail: str = "") -> np.ndarray:
    """GFSK carrier at complex baseband whose bits are random except for `kind`'s header inserted every
    `period_bits` bits (followed by `tail`, e.g. M10/M20 frame-type bits) -> cs16 [nsamp, 2].
    SNR is carrier-to-noise in the full Fs bandwidth."""
    rng = np.random.default_rng(seed)
    _, sps, hdr = HEADER_BANK[kind]
    spb = Fs / sps
    nbits = int(nsamp / spb) + 8
    bits = rng.integers(0, 2, size=nbits, dtype=np.uint8)
    h = np.frombuffer((hdr + tail).encode(), dtype=np.uint8) - ord("0")
    if period_bits is None:
        period_bits = max(4 * len(h), int(0.25 * sps))
    for s in range(int(rng.integers(len(h), period_bits)), nbits - len(h), period_bits):
        bits[s:s + len(h)] = h
    lvl = bits.astype(np.float64) * 2.0 - 1.0
    if invert:
        lvl = -lvl
    t0 = float(rng.uniform(0, 1))
    idx = (np.arange(nsamp, dtype=np.float64) / spb + t0).astype(np.int64)
    f = lvl[idx]
    # Gaussian-ish pulse shaping: moving average over ~0.6 bit (keeps the header eye open at every spb)
    w = max(1, int(round(0.6 * spb)))
    f = np.convolve(f, np.ones(w) / w, mode="same")
    if f_off is None:
        f_off = float(rng.uniform(-3000.0, 3000.0))
    phase = 2.0 * np.pi * np.cumsum(dev_hz * f + f_off) / Fs + float(rng.uniform(0, 2 * np.pi))
    np.remainder(phase, 2.0 * np.pi, out=phase)
    sigma = amp / np.sqrt(2.0) / (10.0 ** (snr_db / 20.0))
    out = rng.standard_normal((nsamp, 2), dtype=np.float32) * np.float32(sigma)
    out[:, 0] += np.float32(amp) * np.cos(phase).astype(np.float32)
    out[:, 1] += np.float32(amp) * np.sin(phase).astype(np.float32)
    np.rint(out, out=out)
    np.clip(out, -32768, 32767, out=out)
    return out.astype(np.int16)


def wideband_mix(seed: int, nsamp: int, sr_base: int, emitters, amp: float = 4000.0, snr_db: float = 20.0,
                 dev_hz: float = 2400.0, dc: tuple[float, float] = (0.0, 0.0)) -> np.ndarray:
    """One wideband capture (cs16 [nsamp, 2] at `sr_base`) holding several GFSK sondes: `emitters` is a list of
    (kind, centre_hz[, tail[, invert]]) with `kind` a HEADER_BANK row.  Each carrier repeats its header like
    fm_header_iq; noise (per-carrier SNR in the full sr_base bandwidth) and an optional DC offset are added once.
    Used for the `dft_detect --IQ <fq>` (translate + decimate) path, SURVEY 8f-2."""
    rng = np.random.default_rng(seed)
    acc = np.zeros((nsamp, 2), dtype=np.float64)
    n = np.arange(nsamp, dtype=np.float64)
    for e in emitters:
        kind, fc = e[0], float(e[1])
        tail = e[2] if len(e) > 2 else ""
        invert = bool(e[3]) if len(e) > 3 else False
        _, sps, hdr = HEADER_BANK[kind]
        spb = sr_base / sps
        nbits = int(nsamp / spb) + 8
        bits = rng.integers(0, 2, size=nbits, dtype=np.uint8)
        h = np.frombuffer((hdr + tail).encode(), dtype=np.uint8) - ord("0")
        period_bits = max(4 * len(h), int(0.25 * sps))
        for s in range(int(rng.integers(len(h), period_bits)), nbits - len(h), period_bits):
            bits[s:s + len(h)] = h
        lvl = bits.astype(np.float64) * 2.0 - 1.0
        if invert:
            lvl = -lvl
        idx = (n / spb + float(rng.uniform(0, 1))).astype(np.int64)
        f = lvl[idx]
        w = max(1, int(round(0.6 * spb)))
        if w > 1:                                        # boxcar via cumulative sum (w can be thousands of samples here)
            cs = np.concatenate(([0.0], np.cumsum(f)))
            lo = np.clip(np.arange(nsamp) - w // 2, 0, nsamp)
            hi = np.clip(np.arange(nsamp) - w // 2 + w, 0, nsamp)
            f = (cs[hi] - cs[lo]) / w
        phase = 2.0 * np.pi * (np.cumsum(dev_hz * f) / sr_base + np.remainder(fc * n / sr_base, 1.0)) + float(rng.uniform(0, 2 * np.pi))
        acc[:, 0] += amp * np.cos(phase)
        acc[:, 1] += amp * np.sin(phase)
    sigma = amp / np.sqrt(2.0) / (10.0 ** (snr_db / 20.0))
    acc += rng.standard_normal((nsamp, 2)) * sigma
    acc[:, 0] += dc[0]
    acc[:, 1] += dc[1]
    np.rint(acc, out=acc)
    np.clip(acc, -32768, 32767, out=acc)
    return acc.astype(np.int16)
