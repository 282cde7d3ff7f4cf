"""Host-side mirror of the reference detector CLI (scan/dft_detect.c) over libb200dft.so.

``DftDetectBatch(sample_rate, bw_khz, dc, nchan)`` is "``nchan`` dft_detect processes with the arguments
``--iq --bw <bw_khz> [--dc] - <sample_rate> <bits>``"; ``run()`` returns for every channel what that process would
have printed (type, score, sample position, frequency offset) and its exit code (+-type number,
dft_detect.c:1660-1666).  ``format_lines()`` reproduces the stdout text parsed by auto_rx
(auto_rx/autorx/scan.py:227-410).  No CPU fallback; nothing under oracle/ is used here.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb200dft.so")
_FMT = {np.dtype(np.int16): 0, np.dtype(np.float32): 1, np.dtype(np.uint8): 2}


class B200DftError(RuntimeError):
    pass


class _Det(C.Structure):
    _fields_ = [("hdr", C.c_int), ("tn", C.c_int), ("score", C.c_float), ("pos", C.c_uint32), ("df_hz", C.c_float),
                ("m10", C.c_uint32), ("type", C.c_char * 12)]


class _Res(C.Structure):
    _fields_ = [("exit_code", C.c_int), ("ndet", C.c_int), ("dets", _Det * 64)]


class _Info(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("sample_rate", "N_DFT", "K", "L", "M", "delay", "lpIQtaps", "lpFMtaps", "nhdr_active", "nchan")]


class _WbInfo(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("sr_base", "if_sr", "decM", "dectaps", "nchan")]


@dataclass
class Detection:
    type: str
    score: float
    pos: int
    df_hz: float
    tn: int
    m10: int


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise B200DftError(f"{LIB_PATH} not built (run __graft_entry__.build()); there is no CPU fallback")
        L = C.CDLL(LIB_PATH)
        L.b200dft_create.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_double, C.c_int, C.c_int, C.c_int]
        L.b200dft_destroy.argtypes = [C.c_void_p]
        L.b200dft_destroy.restype = None
        L.b200dft_info.argtypes = [C.c_void_p, C.POINTER(_Info)]
        L.b200dft_last_error.restype = C.c_char_p
        L.b200dft_run_host.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_long, C.c_long, C.c_int, C.c_double, C.c_void_p]
        L.b200dft_run_device.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_long, C.c_long, C.c_int, C.c_double, C.c_void_p]
        L.b200dft_stream_begin.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double]
        L.b200dft_stream_push_host.argtypes = [C.c_void_p, C.c_void_p, C.c_long, C.c_long, C.c_int, C.c_void_p, C.c_void_p]
        L.b200dft_launch_count.argtypes = [C.c_void_p]
        L.b200dft_launch_count.restype = C.c_long
        L.b200dft_last_kernel_ms.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float)]
        L.b200dft_create_wideband.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_double, C.c_int, C.c_int, C.c_int]
        L.b200dft_wideband_info.argtypes = [C.c_void_p, C.POINTER(_WbInfo)]
        L.b200dft_run_wideband_host.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_long, C.c_void_p, C.c_int, C.c_double, C.c_void_p]
        L.b200dft_run_wideband_device.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_long, C.c_void_p, C.c_int, C.c_double, C.c_void_p]
        L.b200dft_channelise_device.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_long, C.c_void_p, C.c_void_p, C.c_long, C.POINTER(C.c_long)]
        L.b200dft_last_channelise_ms.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
        _lib = L
    return _lib


def _ck(rc: int) -> None:
    if rc != 0:
        raise B200DftError(f"libb200dft error {rc}: {lib().b200dft_last_error().decode()}")


class DftDetectBatch:
    def __init__(self, sample_rate: int = 48000, bw_khz: float = 15.0, dc: bool = True, nchan: int = 1, device: int = 0):
        self.L = lib()
        self.h = C.c_void_p()
        _ck(self.L.b200dft_create(C.byref(self.h), sample_rate, float(bw_khz), int(dc), nchan, device))
        info = _Info()
        _ck(self.L.b200dft_info(self.h, C.byref(info)))
        self.info = info
        self.nchan, self.sample_rate, self.dc = nchan, sample_rate, dc
        self.K, self.N_DFT = info.K, info.N_DFT

    def close(self):
        if getattr(self, "h", None):
            self.L.b200dft_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _collect(self, res):
        out = []
        for c in range(self.nchan):
            r = res[c]
            dets = [Detection(r.dets[i].type.decode(), r.dets[i].score, r.dets[i].pos, r.dets[i].df_hz, r.dets[i].tn, r.dets[i].m10)
                    for i in range(min(r.ndet, 64))]
            out.append((r.exit_code, dets))
        return out

    def run(self, iq: np.ndarray, cont: bool = False, tl: float = -1.0):
        """iq: [nchan, nsamp, 2] int16 / float32 / uint8 host array -> [(exit_code, [Detection...])] per channel."""
        iq = np.ascontiguousarray(iq)
        fmt = _FMT[iq.dtype]
        iq = iq.reshape(self.nchan, -1)
        nsamp = iq.shape[1] // 2
        res = (_Res * self.nchan)()
        _ck(self.L.b200dft_run_host(self.h, iq.ctypes.data, fmt, nsamp, nsamp, int(cont), float(tl), res))
        return self._collect(res)

    def run_device(self, d_iq, fmt: str, nsamp: int, cont: bool = False, tl: float = -1.0, in_stride: int | None = None):
        f = {"cs16": 0, "f32": 1, "cu8": 2}[fmt]
        res = (_Res * self.nchan)()
        ptr = d_iq.data_ptr() if hasattr(d_iq, "data_ptr") else int(d_iq)
        _ck(self.L.b200dft_run_device(self.h, C.c_void_p(ptr), f, in_stride or nsamp, nsamp, int(cont), float(tl), res))
        return self._collect(res)

    # ---- streaming form: what one long-running dft_detect process does between reads (include/b200dft.h) ----
    def stream_begin(self, fmt: str = "cs16", cont: bool = False, tl: float = -1.0) -> None:
        self._sres = (_Res * self.nchan)()
        self._sdone = (C.c_int * self.nchan)()
        self._sdets = [[] for _ in range(self.nchan)]
        self._sfmt = fmt
        _ck(self.L.b200dft_stream_begin(self.h, {"cs16": 0, "f32": 1, "cu8": 2}[fmt], int(cont), float(tl)))

    def stream_push(self, iq: np.ndarray, flush: bool = False):
        """append iq [nchan, n, 2] to the streams -> [(exit_code or None while running, all detections so far)] per channel"""
        iq = np.ascontiguousarray(iq).reshape(self.nchan, -1)
        n = iq.shape[1] // 2
        _ck(self.L.b200dft_stream_push_host(self.h, iq.ctypes.data if n else None, n, n, int(flush), self._sres, self._sdone))
        out = []
        for c in range(self.nchan):
            r = self._sres[c]
            assert r.ndet <= 64, "more than B200DFT_MAXDET detections in one push"
            self._sdets[c] += [Detection(r.dets[i].type.decode(), r.dets[i].score, r.dets[i].pos, r.dets[i].df_hz, r.dets[i].tn, r.dets[i].m10)
                               for i in range(r.ndet)]
            r.ndet = 0                                    # printed: the library appends the next ones from slot 0
            out.append((r.exit_code if self._sdone[c] else None, list(self._sdets[c])))
        return out

    def kernel_ms(self):
        a, b = C.c_float(0), C.c_float(0)
        _ck(self.L.b200dft_last_kernel_ms(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    @property
    def launches(self) -> int:
        return int(self.L.b200dft_launch_count(self.h))

    def format_lines(self, dets, verbose: bool = False) -> str:
        """The stdout text of the reference (dft_detect.c:1620-1634)."""
        out = []
        for d in dets:
            if verbose:
                out.append(f"sample: {d.pos}")
            ln = f"{d.type}: {d.score:.4f}"
            if d.type.startswith(("M10", "M20")) and verbose:
                ln += f" [{d.m10 & 0xFFFF:04X}]"
            if self.dc:
                ln += f" , {d.df_hz:+.1f}Hz"
            out.append(ln)
        return "\n".join(out) + ("\n" if out else "")


class DftWidebandBatch(DftDetectBatch):
    """``nchan`` x ``dft_detect --IQ <fq> --bw <bw_khz> [--dc] - <sr_base> <bits>`` over ONE wideband capture
    (dft_detect.c:1383-1392: rotate by -fq*sr_base, decimate to the IF rate, detect).  ``fq`` is the reference's
    argument: the candidate's offset as a fraction of ``sr_base``."""

    def __init__(self, sr_base: int, bw_khz: float = 15.0, dc: bool = True, nchan: int = 1, device: int = 0):
        self.L = lib()
        self.h = C.c_void_p()
        _ck(self.L.b200dft_create_wideband(C.byref(self.h), sr_base, float(bw_khz), int(dc), nchan, device))
        info, wb = _Info(), _WbInfo()
        _ck(self.L.b200dft_info(self.h, C.byref(info)))
        _ck(self.L.b200dft_wideband_info(self.h, C.byref(wb)))
        self.info, self.wb = info, wb
        self.nchan, self.sr_base, self.sample_rate, self.dc = nchan, sr_base, wb.if_sr, dc
        self.decM, self.dectaps = wb.decM, wb.dectaps
        self.K, self.N_DFT = info.K, info.N_DFT

    def _fq(self, fq):
        fq = np.ascontiguousarray(fq, dtype=np.float64)
        if fq.shape != (self.nchan,):
            raise B200DftError(f"fq must hold {self.nchan} values")
        return fq

    def run(self, wide: np.ndarray, fq, cont: bool = False, tl: float = -1.0):
        """wide: [nwide, 2] int16 / float32 / uint8 host array -> [(exit_code, [Detection...])] per fq."""
        wide = np.ascontiguousarray(wide)
        fq = self._fq(fq)
        res = (_Res * self.nchan)()
        _ck(self.L.b200dft_run_wideband_host(self.h, wide.ctypes.data, _FMT[wide.dtype], wide.size // 2, fq.ctypes.data, int(cont), float(tl), res))
        return self._collect(res)

    def run_device(self, d_wide, fmt: str, nwide: int, fq, cont: bool = False, tl: float = -1.0):
        fq = self._fq(fq)
        res = (_Res * self.nchan)()
        ptr = d_wide.data_ptr() if hasattr(d_wide, "data_ptr") else int(d_wide)
        _ck(self.L.b200dft_run_wideband_device(self.h, C.c_void_p(ptr), {"cs16": 0, "f32": 1, "cu8": 2}[fmt], nwide, fq.ctypes.data,
                                               int(cont), float(tl), res))
        return self._collect(res)

    def channelise(self, d_wide, fmt: str, nwide: int, fq, d_out, out_stride: int | None = None) -> int:
        """IF streams only: d_out (device, float32 [nchan, out_stride, 2]) <- channel c's decimated samples; returns nout."""
        fq = self._fq(fq)
        n = C.c_long(0)
        ptr = d_wide.data_ptr() if hasattr(d_wide, "data_ptr") else int(d_wide)
        optr = d_out.data_ptr() if hasattr(d_out, "data_ptr") else int(d_out)
        _ck(self.L.b200dft_channelise_device(self.h, C.c_void_p(ptr), {"cs16": 0, "f32": 1, "cu8": 2}[fmt], nwide, fq.ctypes.data,
                                             C.c_void_p(optr), out_stride or nwide // self.decM, C.byref(n)))
        return n.value

    def channelise_ms(self) -> float:
        ms = C.c_float(0)
        _ck(self.L.b200dft_last_channelise_ms(self.h, C.byref(ms)))
        return ms.value


# ---- scanner peak stage ------------------------------------------------------------------------------------
def scan_psd(d_wide, fmt: str, nwide: int, sr_base: int, centre_hz: float = 0.0, nfft: int = 8192, device: int = 0):
    """(freq_hz, power_db, step_hz, kernel_ms): the triple auto_rx reads from rtl_power (sdr_wrappers.get_power_spectrum), computed on
    the GPU from the capture the channeliser uses."""
    L = lib()
    L.b200dft_scan_psd_device.argtypes = [C.c_void_p, C.c_int, C.c_long, C.c_int, C.c_double, C.c_int, C.c_void_p, C.c_void_p,
                                          C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_float)]
    L.b200dft_scan_last_error.restype = C.c_char_p
    f, p = np.zeros(nfft), np.zeros(nfft)
    step, ms = C.c_double(0), C.c_float(0)
    ptr = d_wide.data_ptr() if hasattr(d_wide, "data_ptr") else int(d_wide)
    rc = L.b200dft_scan_psd_device(C.c_void_p(ptr), {"cs16": 0, "f32": 1, "cu8": 2}[fmt], nwide, sr_base, float(centre_hz), nfft,
                                   f.ctypes.data, p.ctypes.data, C.byref(step), device, C.byref(ms))
    if rc:
        raise B200DftError(f"scan_psd error {rc}: {L.b200dft_scan_last_error().decode()}")
    return f, p, step.value, ms.value


def detect_peaks(x, mph=None, mpd=1):
    """autorx.utils.detect_peaks(x, mph, mpd) (utils.py:435) -> ascending peak indices."""
    L = lib()
    L.b200dft_scan_detect_peaks.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_void_p, C.c_int]
    x = np.ascontiguousarray(x, dtype=np.float64)
    idx = np.zeros(max(x.size, 1), np.int32)
    n = L.b200dft_scan_detect_peaks(x.ctypes.data, x.size, int(mph is not None), float(mph or 0.0), float(mpd), idx.ctypes.data, idx.size)
    if n < 0:
        raise B200DftError(f"detect_peaks error {n}")
    return idx[:n].astype(np.int64)


def scan_candidates(freq_hz, power_db, step_hz, snr_threshold=10.0, min_distance=1000.0, quantization=10000.0, min_freq_hz=-np.inf,
                    max_freq_hz=np.inf, max_peaks=10):
    """The candidate list of SondeScanner.sonde_search (scan.py:1005-1063) -> (frequencies in Hz strongest first, noise floor dB)."""
    L = lib()
    L.b200dft_scan_candidates.argtypes = [C.c_void_p, C.c_void_p, C.c_int] + [C.c_double] * 6 + [C.c_int, C.c_void_p, C.c_int, C.POINTER(C.c_double)]
    f = np.ascontiguousarray(freq_hz, dtype=np.float64)
    p = np.ascontiguousarray(power_db, dtype=np.float64)
    out = np.zeros(max(f.size, 1))
    nf = C.c_double(0)
    lo = -1e300 if not np.isfinite(min_freq_hz) else float(min_freq_hz)
    hi = 1e300 if not np.isfinite(max_freq_hz) else float(max_freq_hz)
    n = L.b200dft_scan_candidates(f.ctypes.data, p.ctypes.data, f.size, float(step_hz), float(snr_threshold), float(min_distance), float(quantization),
                                  lo, hi, int(max_peaks), out.ctypes.data, out.size, C.byref(nf))
    if n < 0:
        raise B200DftError(f"scan_candidates error {n}")
    return out[:n].copy(), nf.value
