"""Test helper: well-formed RS41 frames (valid block CRCs + Reed-Solomon parity) as a soft-bit stream, the input format of
`rs41mod --softin -i` (rs41mod.c:2875-2960).  Uses the oracle's encoder/CRC, which tests/test_ecc.py pins against the reference
module; a wrong encoder would also show as "[NO]" on clean frames in the reference's own output."""
import numpy as np

from oracle import pyoracle as po

HDR = "0000100001101101010100111000100001000100011010010100100000011111"          # rs41mod.c:170
HDR_BYTES = bytes([0x86, 0x35, 0xf4, 0x40, 0x93, 0xdf, 0x1a, 0x60])
MASK = bytes([0x96, 0x83, 0x3E, 0x51, 0xB1, 0x49, 0x08, 0x98, 0x32, 0x05, 0x59, 0x0E, 0xF9, 0x44, 0xC6, 0x26,
              0x21, 0x60, 0xC2, 0xEA, 0x79, 0x5D, 0x6D, 0xA1, 0x54, 0x69, 0x47, 0x0C, 0xDC, 0xE8, 0x5C, 0xF1,
              0xF7, 0x76, 0x82, 0x7F, 0x07, 0x99, 0xA2, 0x2C, 0x93, 0x7C, 0x30, 0x63, 0xF5, 0x10, 0x2E, 0x61,
              0xD0, 0xBC, 0xB4, 0xB6, 0x06, 0xAA, 0xF4, 0x23, 0x78, 0x6E, 0x3B, 0xAE, 0xBF, 0x7B, 0x4C, 0xC1])
BLOCKS = [(0x039, 0x79, 0x28), (0x065, 0x7A, 0x2A), (0x093, 0x7C, 0x1E), (0x0B5, 0x7D, 0x59), (0x112, 0x7B, 0x15), (0x12B, 0x76, 0x11)]


def clean_frame(rng: np.random.Generator, frnb: int = 1) -> np.ndarray:
    """320-byte standard frame: header, 48 parity bytes, type byte 0x0F, six CRC-protected blocks (random payload)."""
    f = np.zeros(518, np.uint8)
    f[:8] = np.frombuffer(HDR_BYTES, np.uint8)
    f[0x38] = 0x0F
    for pos, typ, ln in BLOCKS:
        f[pos], f[pos + 1] = typ, ln
        if typ != 0x76:
            f[pos + 2: pos + 2 + ln] = rng.integers(0, 256, ln, dtype=np.uint8)
        crc = po.crc16(f[pos + 2: pos + 2 + ln].tobytes())
        f[pos + 2 + ln], f[pos + 3 + ln] = crc & 0xFF, crc >> 8
    f[0x3B], f[0x3C] = frnb & 0xFF, (frnb >> 8) & 0xFF
    crc = po.crc16(f[0x3B: 0x3B + 0x28].tobytes())
    f[0x3B + 0x28], f[0x3C + 0x28] = crc & 0xFF, crc >> 8
    cw1 = po.rs255_encode(f[56:518:2])
    cw2 = po.rs255_encode(f[57:518:2])
    f[8:32], f[32:56] = cw1[:24], cw2[:24]
    return f[:320].copy()


def frame_bits(frame: np.ndarray) -> np.ndarray:
    """header bit string + whitened payload bytes 8.., little-endian bits (rs41mod.c:224-234, 175-182)"""
    hb = np.frombuffer(HDR.encode(), np.uint8) - 48
    pay = frame[8:] ^ np.frombuffer(bytes(MASK[i % 64] for i in range(8, frame.size)), np.uint8)
    bits = np.unpackbits(pay, bitorder="little")
    return np.concatenate([hb, bits]).astype(np.uint8)


def soft_stream(seed: int, nframes: int, sigma: float, gap_bits: int = 2240, burst=None, amp: float = 1.0):
    """-> (float32 soft bits for `--softin -i`, list of clean frames).  Each frame is followed by `gap_bits` random bits (the
    reference always reads 510 bytes after a header).  burst = (frame index, first byte, count): those bytes get random bits."""
    rng = np.random.default_rng(seed)
    out, frames = [rng.integers(0, 2, 200, dtype=np.uint8)], []
    for k in range(nframes):
        fr = clean_frame(rng, frnb=k + 1)
        frames.append(fr)
        b = frame_bits(fr)
        if burst is not None and burst[0] == k:
            lo = 64 + 8 * (burst[1] - 8)
            b[lo: lo + 8 * burst[2]] = rng.integers(0, 2, 8 * burst[2], dtype=np.uint8)
        out += [b, rng.integers(0, 2, gap_bits, dtype=np.uint8)]
    bits = np.concatenate(out)
    soft = -(2.0 * bits - 1.0) * amp                       # -i: a transmitted 1 arrives negative (rs41mod.c:2925-2930)
    soft = soft + rng.standard_normal(soft.size) * sigma
    return soft.astype(np.float32), frames
