"""ctypes access to the oracle (TEST INFRASTRUCTURE ONLY -- see oracle/oracle.h).

  * ``liboracle.so``          our plain-C restatement (oracle/fsk_oracle.c, dft_oracle.c)
  * ``_ref/libfsk_ref.so``    the UNMODIFIED reference modem + our in-memory frame pump
  * ``_ref/fsk_demod``, ``_ref/dft_detect``   the reference CLIs

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
FMT = {"cs16": 0, "f32": 1, "cu8": 2, "s16": 3}


def build(verbose: bool = False) -> None:
    """Compile liboracle.so and (when /root/reference is present) oracle/_ref/."""
    r = subprocess.run(["make", "-C", HERE, "all"], capture_output=True, text=True)
    if verbose or r.returncode:
        print(r.stdout, r.stderr)
    if r.returncode:
        raise RuntimeError("oracle build failed")


def _load(path: str) -> C.CDLL:
    if not os.path.exists(path):
        build()
    return C.CDLL(path)


class FrameRec(C.Structure):
    _fields_ = [("nin_used", C.c_int), ("nin_next", C.c_int), ("norm_rx_timing", C.c_float),
                ("rx_timing", C.c_float), ("ppm", C.c_float), ("EbNodB", C.c_float), ("snr_est", C.c_float),
                ("foff", C.c_float), ("f_est", C.c_float * 4), ("mean_sum", C.c_float), ("std_sum", C.c_float)]


REC_DTYPE = np.dtype([("nin_used", "i4"), ("nin_next", "i4"), ("norm_rx_timing", "f4"), ("rx_timing", "f4"),
                      ("ppm", "f4"), ("EbNodB", "f4"), ("snr_est", "f4"), ("foff", "f4"), ("f_est", "f4", (4,)),
                      ("mean_sum", "f4"), ("std_sum", "f4")])
assert REC_DTYPE.itemsize == C.sizeof(FrameRec)

_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        L = _load(os.path.join(HERE, "liboracle.so"))
        L.fsko_create.restype = C.c_void_p
        L.fsko_create.argtypes = [C.c_int] * 7
        L.fsko_destroy.argtypes = [C.c_void_p]
        L.fsko_set_freq_est_limits.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.fsko_set_freq_est_alg.argtypes = [C.c_void_p, C.c_int]
        L.fsko_nin.argtypes = [C.c_void_p]
        L.fsko_nbits.argtypes = [C.c_void_p]
        L.fsko_ndft.argtypes = [C.c_void_p]
        L.fsko_param.argtypes = [C.c_void_p, C.c_char_p]
        L.fsko_frame.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        for f in ("fsko_Sf", "fsko_hann", "fsko_twiddles", "fsko_f_int", "fsko_f_dc", "fsko_phi_c"):
            getattr(L, f).restype = C.POINTER(C.c_float)
            getattr(L, f).argtypes = [C.c_void_p]
        L.fsko_last_frame.argtypes = [C.c_void_p, C.POINTER(FrameRec)]
        L.fsko_fft.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.fsko_eye.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_int)]
        L.fsko_run_stream.restype = C.c_long
        L.fsko_run_stream.argtypes = [C.c_void_p, C.c_void_p, C.c_long, C.c_int, C.c_int, C.c_int,
                                      C.c_void_p, C.POINTER(C.c_long), C.c_void_p, C.c_long]
        _lib = L
    return _lib


@dataclass
class FskConfig:
    """fsk_demod command-line parameters (utils/fsk_demod.c:53-232)."""
    Fs: int = 48000
    Rs: int = 4800
    M: int = 2
    P: int = 10                 # CLI default (fsk_demod.c:71)
    Nsym: int = 50              # FSK_DEFAULT_NSYM (fsk.h:46)
    lower: int | None = None    # --fsk_lower; None -> -Fs/2 for complex input (fsk_demod.c:218-226)
    upper: int | None = None
    mask: int = 0               # --mask TxSpacing: 0 = peak estimator
    soft: bool = True
    softinv: bool = False

    @property
    def fs_tx(self) -> int:
        return self.mask if self.mask else 100           # fsk_demod.c:79,159

    @property
    def lo(self) -> int:
        return -self.Fs // 2 if self.lower is None else self.lower

    @property
    def hi(self) -> int:
        return self.Fs // 2 if self.upper is None else self.upper

    def cli_args(self, fmt: str = "cs16") -> list[str]:
        a = {"cs16": ["--cs16"], "cu8": ["--cu8"], "s16": []}[fmt]
        if self.lower is not None:
            a += ["-b", str(self.lower)]
        if self.upper is not None:
            a += ["-u", str(self.upper)]
        if self.soft:
            a += ["-s"]
        if self.softinv:
            a += ["-i"]
        if self.mask:
            a += ["--mask", str(self.mask)]
        a += [f"--nsym={self.Nsym}", "-p", str(self.P), str(self.M), str(self.Fs), str(self.Rs)]
        return a


RS41 = FskConfig(48000, 4800, 2, 5, 300, -5000, 5000, 5000)          # auto_rx/autorx/decode.py:901


class Oracle:
    """One channel of the C restatement."""

    def __init__(self, cfg: FskConfig):
        self.cfg = cfg
        self.L = lib()
        self.h = self.L.fsko_create(cfg.Fs, cfg.Rs, cfg.M, cfg.P, cfg.Nsym, 1000, cfg.fs_tx)
        self.L.fsko_set_freq_est_limits(self.h, cfg.lo, cfg.hi)
        self.L.fsko_set_freq_est_alg(self.h, 1 if cfg.mask else 0)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.fsko_destroy(self.h)
            self.h = None

    def param(self, k: str) -> int:
        return self.L.fsko_param(self.h, k.encode())

    @property
    def nin(self) -> int:
        return self.L.fsko_nin(self.h)

    def _arr(self, fn: str, n: int) -> np.ndarray:
        return np.ctypeslib.as_array(getattr(self.L, fn)(self.h), shape=(n,)).copy()

    def Sf(self):
        return self._arr("fsko_Sf", self.param("Ndft"))

    def hann(self):
        return self._arr("fsko_hann", self.param("Ndft"))

    def twiddles(self):
        return self._arr("fsko_twiddles", 2 * self.param("Ndft"))

    def f_int(self):
        return self._arr("fsko_f_int", 2 * self.param("M") * (self.param("Nsym") + 1) * self.param("P"))

    def f_dc(self):
        return self._arr("fsko_f_dc", 2 * self.param("M") * self.param("Nmem"))

    def phi_c(self):
        return self._arr("fsko_phi_c", 2 * self.param("M"))

    def fft(self, x: np.ndarray) -> np.ndarray:
        x = np.ascontiguousarray(x, dtype=np.float32)
        y = np.empty_like(x)
        self.L.fsko_fft(self.h, x.ctypes.data, y.ctypes.data)
        return y

    def frame(self, iq: np.ndarray, soft: bool = True):
        iq = np.ascontiguousarray(iq, dtype=np.float32)
        nb = self.L.fsko_nbits(self.h)
        if soft:
            o = np.zeros(nb, np.float32)
            self.L.fsko_frame(self.h, iq.ctypes.data, None, o.ctypes.data)
        else:
            o = np.zeros(nb, np.uint8)
            self.L.fsko_frame(self.h, iq.ctypes.data, o.ctypes.data, None)
        return o

    def last(self) -> FrameRec:
        r = FrameRec()
        self.L.fsko_last_frame(self.h, C.byref(r))
        return r

    def run_stream(self, data: np.ndarray, fmt: str = "cs16"):
        """CLI frame pump over memory -> (out, consumed, records)."""
        data = np.ascontiguousarray(data)
        per = {"cs16": 2, "f32": 2, "cu8": 2, "s16": 1}[fmt]
        nsamp = data.size // per
        nb = self.L.fsko_nbits(self.h)
        maxf = nsamp // max(1, self.param("N") - self.param("Ts") // 2) + 2
        out = np.zeros(maxf * nb, np.float32 if self.cfg.soft else np.uint8)
        recs = np.zeros(maxf, REC_DTYPE)
        cons = C.c_long(0)
        nf = self.L.fsko_run_stream(self.h, data.ctypes.data, nsamp, FMT[fmt], int(self.cfg.soft),
                                    int(self.cfg.softinv), out.ctypes.data, C.byref(cons), recs.ctypes.data, maxf)
        return out[:nf * nb], cons.value, recs[:nf]


# ------------------------------------------------------------------------------------------------
# the UNMODIFIED reference (oracle/_ref)
# ------------------------------------------------------------------------------------------------
_ref = None
REF_REC = 10


def ref_available() -> bool:
    return os.path.exists(os.path.join(REF_DIR, "libfsk_ref.so")) or os.path.isdir("/root/reference/utils")


def reflib() -> C.CDLL:
    global _ref
    if _ref is None:
        L = _load(os.path.join(REF_DIR, "libfsk_ref.so"))
        L.ref_fsk_run_stream.restype = C.c_long
        L.ref_fsk_run_stream.argtypes = [C.c_int] * 9 + [C.c_void_p, C.c_long, C.c_int, C.c_int, C.c_int,
                                                         C.c_void_p, C.POINTER(C.c_long), C.c_void_p, C.c_long, C.c_void_p]
        L.ref_fsk_run_stream_passes.restype = C.c_long
        L.ref_fsk_run_stream_passes.argtypes = L.ref_fsk_run_stream.argtypes + [C.c_int]
        _ref = L
    return _ref


def ref_run_stream(cfg: FskConfig, data: np.ndarray, fmt: str = "cs16", want_out: bool = True):
    """Reference modem over an in-memory stream -> (out, consumed, recs[nf,10], Sf)."""
    L = reflib()
    data = np.ascontiguousarray(data)
    per = {"cs16": 2, "f32": 2, "cu8": 2, "s16": 1}[fmt]
    nsamp = data.size // per
    Ts = cfg.Fs // cfg.Rs
    N = Ts * cfg.Nsym
    nb = cfg.Nsym * (1 if cfg.M == 2 else 2)
    maxf = nsamp // max(1, N - Ts // 2) + 2
    out = np.zeros(maxf * nb, np.float32 if cfg.soft else np.uint8) if want_out else None
    recs = np.zeros((maxf, REF_REC), np.float32)
    Sf = np.zeros(4096, np.float32)
    cons = C.c_long(0)
    nf = L.ref_fsk_run_stream(cfg.Fs, cfg.Rs, cfg.M, cfg.P, cfg.Nsym, cfg.fs_tx, cfg.lo, cfg.hi, 1 if cfg.mask else 0,
                              data.ctypes.data, nsamp, FMT[fmt], int(cfg.soft), int(cfg.softinv),
                              out.ctypes.data if want_out else None, C.byref(cons), recs.ctypes.data, maxf,
                              Sf.ctypes.data)
    return (out[:nf * nb] if want_out else None), cons.value, recs[:nf], Sf


def ref_cli_fsk_demod(cfg: FskConfig, raw: bytes, fmt: str = "cs16", extra: list[str] | None = None):
    """Run the reference fsk_demod binary on stdin bytes -> (stdout bytes, stderr bytes, rc)."""
    exe = os.path.join(REF_DIR, "fsk_demod")
    if not os.path.exists(exe):
        build()
    args = [exe] + (extra or []) + cfg.cli_args(fmt) + ["-", "-"]
    r = subprocess.run(args, input=raw, capture_output=True)
    return r.stdout, r.stderr, r.returncode


# ------------------------------------------------------------------------------------------------
# dft_detect: our restatement (liboracle.so) and the reference binary (oracle/_ref/dft_detect)
# ------------------------------------------------------------------------------------------------
class DftDet(C.Structure):
    _fields_ = [("hdr", C.c_int), ("tn", C.c_int), ("score", C.c_float), ("pos", C.c_uint), ("df_hz", C.c_float),
                ("m10", C.c_uint), ("type", C.c_char * 12)]


class DftBlk(C.Structure):
    _fields_ = [("hdr", C.c_int), ("mp", C.c_int), ("score", C.c_float), ("pos", C.c_uint), ("dc", C.c_float),
                ("sample_in", C.c_uint)]


def dft_oracle_run(iq: np.ndarray, sample_rate: int = 48000, bw_khz: float = 15.0, dc: bool = True, cont: bool = True,
                   tl: float = -1.0, want_blocks: bool = False):
    """-> (exit_code, [(type, score, pos, df_hz)], blocks or None)"""
    L = lib()
    L.dfto_run.restype = C.c_int
    L.dfto_run.argtypes = [C.c_void_p, C.c_long, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, C.c_double,
                           C.c_void_p, C.c_int, C.POINTER(C.c_int), C.c_void_p, C.c_int, C.POINTER(C.c_int)]
    iq = np.ascontiguousarray(iq)
    bits = {np.dtype(np.int16): 16, np.dtype(np.float32): 32, np.dtype(np.uint8): 8}[iq.dtype]
    nsamp = iq.size // 2
    dets = (DftDet * 4096)()
    nd, nb = C.c_int(0), C.c_int(0)
    maxb = (nsamp // 5000 + 2) * 16 if want_blocks else 0
    blks = (DftBlk * max(maxb, 1))()
    rc = L.dfto_run(iq.ctypes.data, nsamp, bits, sample_rate, bw_khz, int(dc), int(cont), tl, dets, 4096, C.byref(nd),
                    blks if want_blocks else None, maxb, C.byref(nb))
    out = [(dets[i].type.decode(), dets[i].score, dets[i].pos, dets[i].df_hz) for i in range(min(nd.value, 4096))]
    bl = [(b.hdr, b.mp, b.score, b.pos, b.dc, b.sample_in) for b in blks[: nb.value]] if want_blocks else None
    return rc, out, bl


def dft_oracle_run_wideband(iq: np.ndarray, sr_base: int, fq: float, bw_khz: float = 15.0, dc: bool = True, cont: bool = True,
                            tl: float = -1.0, want_if: bool = False):
    """`dft_detect --IQ fq ...` restatement -> (exit_code, dets, info dict, IF stream complex64 or None)"""
    L = lib()
    L.dfto_run_wideband.restype = C.c_int
    L.dfto_run_wideband.argtypes = [C.c_void_p, C.c_long, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, C.c_int, C.c_double,
                                    C.c_void_p, C.c_int, C.POINTER(C.c_int), C.c_void_p, C.c_long, C.POINTER(C.c_long),
                                    C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    iq = np.ascontiguousarray(iq)
    bits = {np.dtype(np.int16): 16, np.dtype(np.float32): 32, np.dtype(np.uint8): 8}[iq.dtype]
    nsamp = iq.size // 2
    dets = (DftDet * 4096)()
    nd, ifsr, decm, taps, ifn = C.c_int(0), C.c_int(0), C.c_int(0), C.c_int(0), C.c_long(0)
    ifbuf = np.zeros(nsamp + 8, dtype=np.complex64) if want_if else None
    rc = L.dfto_run_wideband(iq.ctypes.data, nsamp, bits, sr_base, fq, bw_khz, int(dc), int(cont), tl, dets, 4096, C.byref(nd),
                             ifbuf.ctypes.data if want_if else None, nsamp + 8 if want_if else 0, C.byref(ifn),
                             C.byref(ifsr), C.byref(decm), C.byref(taps))
    out = [(dets[i].type.decode(), dets[i].score, dets[i].pos, dets[i].df_hz) for i in range(min(nd.value, 4096))]
    info = {"if_sr": ifsr.value, "decM": decm.value, "dectaps": taps.value, "if_n": ifn.value}
    return rc, out, info, (ifbuf[: ifn.value] if want_if else None)


def ref_cli_dft_detect(iq: np.ndarray, sample_rate: int = 48000, bw_khz: float = 15.0, dc: bool = True, cont: bool = True,
                       tl: float | None = None, bits: int = 16, fq: float | None = None):
    """Run oracle/_ref/dft_detect -v on stdin bytes -> (exit_code as signed int8-ish, [(type, score, pos, df_hz)]).
    fq given: the baseband translate+decimate mode `--IQ fq` at sample_rate = sr_base."""
    exe = os.path.join(REF_DIR, "dft_detect")
    if not os.path.exists(exe):
        build()
    args = [exe, "-v"] + (["-c"] if cont else []) + (["-t", str(tl)] if tl is not None else []) + (["--iq"] if fq is None else ["--IQ", repr(float(fq))]) + ["--bw", str(bw_khz)]
    args += (["--dc"] if dc else []) + ["-", str(sample_rate), str(bits)]
    r = subprocess.run(args, input=np.ascontiguousarray(iq).tobytes(), capture_output=True)
    import re
    dets, pos = [], None
    for ln in r.stdout.decode().splitlines():
        if ln.startswith("sample:"):
            pos = int(ln.split(":")[1])
            continue
        m = re.match(r"^(\w+): (-?[0-9.]+)", ln)
        if m:
            mh = re.search(r", ([+-][0-9.]+)Hz", ln)
            dets.append((m.group(1), float(m.group(2)), pos, float(mh.group(1)) if mh else 0.0))
    rc = r.returncode if r.returncode < 128 else r.returncode - 256
    return rc, dets


# ------------------------------------------------------------------------------------------------
# soft-bit header sync + RS41 frame assembly
# ------------------------------------------------------------------------------------------------
class SsoFrame(C.Structure):
    _fields_ = [("len", C.c_int), ("nbytes_read", C.c_int), ("first_bit", C.c_long), ("score", C.c_float),
                ("bytes", C.c_ubyte * 518), ("bytescore", C.c_float * 518), ("ec", C.c_int), ("crc_mask", C.c_int)]


def softsync_oracle(soft: np.ndarray, softinv: bool = False, inv: bool = True, ecc: int = 0):
    """-> [(len, first_bit, score, bytes[len] as numpy uint8)] like `rs41mod --softin -i -r` would print."""
    L = lib()
    L.sso_rs41.restype = C.c_int
    L.sso_rs41.argtypes = [C.c_void_p, C.c_long, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]
    soft = np.ascontiguousarray(soft, dtype=np.float32)
    cap = soft.size // 4000 + 4
    fr = (SsoFrame * cap)()
    n = L.sso_rs41(soft.ctypes.data, soft.size, int(softinv), int(inv), int(ecc), fr, cap)
    return [(f.len, f.first_bit, f.score, np.frombuffer(bytes(f.bytes), np.uint8)[: f.len].copy(), f.nbytes_read, f.ec, f.crc_mask) for f in fr[: min(n, cap)]]


def ref_cli_rs41mod_raw(soft_bytes: bytes, inv: bool = True, ecc: int = 1):
    """Run oracle/_ref/rs41mod -r [--ecc2] --softin [-i] -> [(hex string, tag)] per printed frame; tag = "[OK] (n)" / "[NO] (-+)"."""
    exe = os.path.join(REF_DIR, "rs41mod")
    r = subprocess.run([exe, "-r"] + (["--ecc2"] if ecc == 2 else []) + ["--softin"] + (["-i"] if inv else []), input=soft_bytes, capture_output=True)
    out = []
    for ln in r.stdout.decode().splitlines():
        parts = ln.split(" ", 1)
        out.append((parts[0], parts[1] if len(parts) > 1 else ""))
    return out


# ------------------------------------------------------------------------------------------------
# Reed-Solomon stage of rs41mod
# ------------------------------------------------------------------------------------------------
def _ecc_lib():
    L = lib()
    L.ecco_rs255_encode.argtypes = [C.c_void_p]
    L.ecco_rs255_decode.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.ecco_rs41_ecc.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.ecco_crc16.argtypes = [C.c_void_p, C.c_int]
    L.ecco_rs41_check_crc.argtypes = [C.c_void_p, C.c_int, C.c_int]
    return L


def rs255_encode(msg: np.ndarray) -> np.ndarray:
    """msg: 231 bytes -> codeword of 255 bytes (parity in cw[0:24], message in cw[24:])."""
    cw = np.zeros(255, np.uint8)
    cw[24:] = msg
    _ecc_lib().ecco_rs255_encode(cw.ctypes.data)
    return cw


def rs255_decode(cw: np.ndarray):
    """-> (return code, corrected copy)"""
    c = np.ascontiguousarray(cw, dtype=np.uint8).copy()
    return _ecc_lib().ecco_rs255_decode(c.ctypes.data, None, None), c


def rs41_ecc(frame: np.ndarray, frmlen: int, level: int = 2):
    f = np.zeros(518, np.uint8)
    f[: frame.size] = frame
    return _ecc_lib().ecco_rs41_ecc(f.ctypes.data, frmlen, level), f


def crc16(data: bytes) -> int:
    b = np.frombuffer(bytes(data), np.uint8)
    return _ecc_lib().ecco_crc16(b.ctypes.data, b.size)


_bch = None


def ref_rs255():
    """The reference module itself (oracle/_ref/libbch_ref.so): (encode(cw), decode(cw) -> (rc, cw))."""
    global _bch
    if _bch is None:
        L = _load(os.path.join(REF_DIR, "libbch_ref.so"))
        rs = (C.c_ubyte * 4096)()                         # RS_t: 7 + 255 + GF_t(4+4+1+256+256) bytes, generously
        L.rs_init_RS255(rs)
        _bch = (L, rs)
    L, rs = _bch

    def enc(cw):
        c = np.ascontiguousarray(cw, dtype=np.uint8).copy()
        L.rs_encode(rs, c.ctypes.data_as(C.c_void_p))
        return c

    def dec(cw):
        c = np.ascontiguousarray(cw, dtype=np.uint8).copy()
        ep, ev = (C.c_ubyte * 32)(), (C.c_ubyte * 32)()
        return L.rs_decode(rs, c.ctypes.data_as(C.c_void_p), ep, ev), c
    return enc, dec


# ------------------------------------------------------------------------------------------------
# scanner peak stage: numpy restatements -- test infrastructure only
# ------------------------------------------------------------------------------------------------
def detect_peaks_oracle(x, mph=None, mpd=1):
    """autorx.utils.detect_peaks (utils.py:435-586) with edge='rising', threshold=0, kpsh=False, valley=False; pinned by
    tests/golden/golden_peaks.json (outputs of the reference function itself)."""
    x = np.atleast_1d(x).astype("float64").copy()
    if x.size < 3:
        return np.array([], dtype=int)
    dx = x[1:] - x[:-1]
    nan = np.isnan(x)
    if nan.any():
        x[nan] = np.inf
        dx[np.isnan(dx)] = np.inf
    nxt, prv = np.append(dx, 0.0), np.insert(dx, 0, 0.0)
    ind = np.where((nxt <= 0) & (prv > 0))[0]                                   # :537-539
    if ind.size and nan.any():
        bad = nan.copy(); bad[:-1] |= nan[1:]; bad[1:] |= nan[:-1]              # :543-550
        ind = ind[~bad[ind]]
    if ind.size and ind[0] == 0:
        ind = ind[1:]
    if ind.size and ind[-1] == x.size - 1:
        ind = ind[:-1]
    if ind.size and mph is not None:
        ind = ind[x[ind] >= mph]
    if ind.size and mpd > 1:                                                     # :564-576
        order = ind[np.argsort(x[ind], kind="stable")][::-1]
        dele = np.zeros(order.size, bool)
        for i in range(order.size):
            if not dele[i]:
                dele |= (order >= order[i] - mpd) & (order <= order[i] + mpd)
                dele[i] = False
        ind = np.sort(order[~dele])
    return ind


def scan_candidates_oracle(freq, power, step, snr_threshold=10.0, min_distance=1000.0, quantization=10000.0, min_freq_hz=-np.inf,
                           max_freq_hz=np.inf, max_peaks=10):
    """The candidate list built inside SondeScanner.sonde_search (scan.py:1005-1063), frequencies in Hz."""
    freq, power = np.asarray(freq, float), np.asarray(power, float)
    nf = np.median(power)
    pk = detect_peaks_oracle(power, mph=nf + snr_threshold, mpd=min_distance / step)
    if pk.size == 0:
        return np.array([]), nf
    pf = freq[pk][np.argsort(power[pk], kind="stable")][::-1]
    pf = np.round(pf / quantization) * quantization
    _, first = np.unique(pf, return_index=True)
    pf = pf[np.sort(first)]
    pf = pf[~((pf < min_freq_hz - quantization / 2.0) | (pf > max_freq_hz + quantization / 2.0))]
    return pf[:max_peaks], nf


def psd_oracle(iq: np.ndarray, nfft: int):
    """float64 averaged periodic-Hann periodogram, fftshifted, dB: what b200dft_scan_psd_device computes (no reference: rtl_power is external)."""
    if iq.dtype == np.int16:
        z = (iq[:, 0].astype(np.float64) + 1j * iq[:, 1]) / 32768.0
    elif iq.dtype == np.uint8:
        z = ((iq[:, 0].astype(np.float64) - 128) + 1j * (iq[:, 1].astype(np.float64) - 128)) / 128.0
    else:
        z = iq[:, 0].astype(np.float64) + 1j * iq[:, 1]
    nseg = z.size // nfft
    w = (0.5 - 0.5 * np.cos(2 * np.pi * np.arange(nfft) / nfft)).astype(np.float32).astype(np.float64)
    S = np.abs(np.fft.fft(z[: nseg * nfft].reshape(nseg, nfft) * w, axis=1)) ** 2
    p = S.mean(axis=0) / (w ** 2).sum()
    return 10.0 * np.log10(np.fft.fftshift(p))
