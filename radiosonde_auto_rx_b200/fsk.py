"""Host-side mirror of the reference modem interface (utils/fsk.h:105-205) over libb200fsk.so.

Two layers:

* :class:`FskBatch` -- the batched B200 path: ``nchan`` independent channels, one modem configuration,
  persistent per-channel demodulator state in HBM.  ``push()`` takes host buffers (the bytes the
  reference CLI would read on stdin, utils/fsk_demod.c:281-309); ``demod_device()`` takes device
  pointers / torch CUDA tensors (plumbing only -- torch is used for memory and streams).
* ``fsk_create_hbr`` / ``fsk_nin`` / ``fsk_demod`` / ``fsk_demod_sd`` / ``fsk_get_demod_stats`` ... --
  single-channel functions with the reference's names, argument meaning and frame semantics, so
  tests read like code written against the reference's fsk.h.

There is NO CPU fallback: importing works anywhere, but creating a modem without the compiled CUDA
library and a GPU raises.  Nothing under oracle/ is used here.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B200FSK_LIB", os.path.join(_HERE, "libb200fsk.so"))   # override only for kernel experiments (profiles/)

FMT_CS16, FMT_F32, FMT_CU8, FMT_S16 = 0, 1, 2, 3
_FMT = {"cs16": 0, "f32": 1, "cu8": 2, "s16": 3}
_ESIZE = {0: 4, 1: 8, 2: 2, 3: 2}

MODE_2FSK, MODE_4FSK = 2, 4          # utils/fsk.h:36-37
FSK_DEFAULT_P, FSK_DEFAULT_NSYM = 8, 50   # utils/fsk.h:45-46


class B200Error(RuntimeError):
    pass


class _Cfg(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("Fs", "Rs", "M", "P", "Nsym", "f1_tx", "fs_tx", "est_min", "est_max", "est_type")]


class _Info(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("Ts", "N", "Nmem", "Ndft", "Nbits", "nin_max", "nin_min", "nchan")] + [("smem_bytes", C.c_size_t)]


class _Stats(C.Structure):
    _fields_ = [("snr_est", C.c_float), ("EbNodB", C.c_float), ("ppm", C.c_float), ("rx_timing", C.c_float),
                ("norm_rx_timing", C.c_float), ("foff", C.c_float), ("f_est", C.c_float * 4), ("nin", C.c_int),
                ("neyetr", C.c_int), ("neyesamp", C.c_int), ("rx_eye", (C.c_float * 160) * 8),
                ("frames", C.c_longlong), ("samples", C.c_longlong)]


_lib = None


def lib() -> C.CDLL:
    """Load libb200fsk.so (built in-tree by __graft_entry__.build()).  Fails loudly if missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise B200Error(f"{LIB_PATH} not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                            "(there is no CPU fallback for the demodulator)")
        L = C.CDLL(LIB_PATH)
        L.b200fsk_create.argtypes = [C.POINTER(C.c_void_p), C.POINTER(_Cfg), C.c_int, C.c_int]
        L.b200fsk_destroy.argtypes = [C.c_void_p]
        L.b200fsk_destroy.restype = None
        L.b200fsk_info.argtypes = [C.c_void_p, C.POINTER(_Info)]
        L.b200fsk_last_error.restype = C.c_char_p
        L.b200fsk_nin.argtypes = [C.c_void_p, C.c_void_p]
        L.b200fsk_clear_estimators.argtypes = [C.c_void_p]
        L.b200fsk_max_frames.argtypes = [C.c_void_p, C.c_long]
        L.b200fsk_max_frames.restype = C.c_long
        L.b200fsk_push_frames_bound.argtypes = [C.c_void_p, C.c_long, C.c_int]
        L.b200fsk_push_frames_bound.restype = C.c_long
        L.b200fsk_demod_device.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_long, C.c_void_p, C.c_long, C.c_int,
                                           C.c_void_p, C.c_long, C.c_int, C.c_int, C.c_void_p]
        L.b200fsk_counts.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.b200fsk_push_host.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_long, C.c_long, C.c_int,
                                        C.c_void_p, C.c_long, C.c_int, C.c_int, C.c_void_p]
        L.b200fsk_enable_stats.argtypes = [C.c_void_p, C.c_int]
        L.b200fsk_get_stats.argtypes = [C.c_void_p, C.c_int, C.POINTER(_Stats)]
        L.b200fsk_get_Sf.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.b200fsk_launch_count.argtypes = [C.c_void_p]
        L.b200fsk_launch_count.restype = C.c_long
        L.b200fsk_last_kernel_ms.argtypes = [C.c_void_p]
        L.b200fsk_last_kernel_ms.restype = C.c_float
        L.b200fsk_last_kernel_kind.argtypes = [C.c_void_p]
        _lib = L
    return _lib


def _ck(rc: int) -> None:
    if rc != 0:
        raise B200Error(f"libb200fsk error {rc}: {lib().b200fsk_last_error().decode()}")


@dataclass
class ModemStats:
    """The fields of struct MODEM_STATS / struct FSK the reference CLI prints (fsk_demod.c:359-420)."""
    snr_est: float
    EbNodB: float
    ppm: float
    rx_timing: float
    norm_rx_timing: float
    foff: float
    f_est: list
    nin: int
    rx_eye: np.ndarray
    frames: int
    samples: int


class FskBatch:
    """``nchan`` concurrent fsk demodulators with the configuration of
    ``fsk_create_hbr(Fs,Rs,M,P,Nsym,f1_tx,fs_tx)`` + ``fsk_set_freq_est_limits`` + ``fsk_set_freq_est_alg``."""

    def __init__(self, Fs: int, Rs: int, M: int = 2, P: int = 10, Nsym: int = FSK_DEFAULT_NSYM, f1_tx: int = 1000,
                 fs_tx: int = 100, est_min: int | None = None, est_max: int | None = None, est_type: int = 0,
                 nchan: int = 1, device: int = 0):
        self.L = lib()
        cfg = _Cfg(Fs, Rs, M, P, Nsym, f1_tx, fs_tx, -Fs // 2 if est_min is None else est_min,
                   Fs // 2 if est_max is None else est_max, int(bool(est_type)))
        self.cfg = cfg
        self.h = C.c_void_p()
        _ck(self.L.b200fsk_create(C.byref(self.h), C.byref(cfg), nchan, device))
        info = _Info()
        _ck(self.L.b200fsk_info(self.h, C.byref(info)))
        self.Ts, self.N, self.Nmem, self.Ndft, self.Nbits = info.Ts, info.N, info.Nmem, info.Ndft, info.Nbits
        self.nin_max, self.nin_min, self.nchan, self.smem_bytes = info.nin_max, info.nin_min, info.nchan, info.smem_bytes
        self.M, self.device = M, device

    def close(self) -> None:
        if getattr(self, "h", None) is not None and self.h:
            self.L.b200fsk_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- sizes ------------------------------------------------------------------------------------
    def max_frames(self, nsamp: int) -> int:
        return int(self.L.b200fsk_max_frames(self.h, nsamp))

    def nin(self) -> np.ndarray:
        """fsk_nin() of every channel."""
        out = np.zeros(self.nchan, np.int32)
        _ck(self.L.b200fsk_nin(self.h, out.ctypes.data))
        return out

    def clear_estimators(self) -> None:
        _ck(self.L.b200fsk_clear_estimators(self.h))

    def enable_stats(self, on: bool = True) -> None:
        _ck(self.L.b200fsk_enable_stats(self.h, int(on)))

    # -- host-buffer path (CLI semantics) ------------------------------------------------------------
    def push(self, data: np.ndarray, fmt: str = "cs16", soft: bool = True, softinv: bool = False,
             max_frames: int = 0, out: np.ndarray | None = None):
        """Append ``data`` ([nchan, nsamp(, 2)]) to the channel streams and demodulate every whole frame.

        Returns ``(out, frames)``: out[c, :frames[c]*Nbits] are channel c's bits (uint8) or soft
        decisions (float32) in stream order.  A trailing partial frame is carried to the next push."""
        f = _FMT[fmt]
        data = np.ascontiguousarray(data)
        per = 1 if f == FMT_S16 else 2
        if data.ndim == 1 or (data.ndim == 2 and per == 2 and self.nchan == 1 and data.shape[-1] == 2):
            data = data.reshape(1, -1)
        data = data.reshape(self.nchan, -1)
        want = {0: np.int16, 1: np.float32, 2: np.uint8, 3: np.int16}[f]
        if data.dtype != want:
            raise TypeError(f"fmt {fmt} needs dtype {want}")
        nsamp = data.shape[1] // per
        stride = int(self.L.b200fsk_push_frames_bound(self.h, nsamp, max_frames)) * self.Nbits
        if out is None:
            out = np.zeros((self.nchan, stride), np.float32 if soft else np.uint8)
        frames = np.zeros(self.nchan, np.int32)
        _ck(self.L.b200fsk_push_host(self.h, data.ctypes.data, f, nsamp, nsamp, max_frames, out.ctypes.data,
                                     out.shape[1], int(soft), int(softinv), frames.ctypes.data))
        return out, frames

    # -- device-resident path ---------------------------------------------------------------------------
    def demod_device(self, d_in, fmt: str, nsamp: int, d_out, soft: bool = True, softinv: bool = False,
                     in_stride: int | None = None, out_stride: int | None = None, max_frames: int = 0,
                     stream=None, d_in_off=None) -> None:
        """Asynchronous demodulation of device-resident streams.  ``d_in`` / ``d_out`` are torch CUDA
        tensors (or raw device addresses): d_in [nchan, in_stride] samples of ``fmt``, d_out
        [nchan, out_stride] float32 (soft) or uint8."""
        f = _FMT[fmt]
        pin = d_in.data_ptr() if hasattr(d_in, "data_ptr") else int(d_in)
        pout = d_out.data_ptr() if hasattr(d_out, "data_ptr") else int(d_out)
        if in_stride is None:
            per = 1 if (f == FMT_S16 or getattr(d_in.dtype, "is_complex", False)) else 2
            in_stride = d_in.numel() // self.nchan // per
        if out_stride is None:
            out_stride = d_out.numel() // self.nchan
        st = None
        if stream is not None:
            st = C.c_void_p(stream.cuda_stream if hasattr(stream, "cuda_stream") else int(stream))
        poff = None
        if d_in_off is not None:
            poff = C.c_void_p(d_in_off.data_ptr() if hasattr(d_in_off, "data_ptr") else int(d_in_off))
        _ck(self.L.b200fsk_demod_device(self.h, C.c_void_p(pin), f, in_stride, poff, nsamp, max_frames,
                                        C.c_void_p(pout), out_stride, int(soft), int(softinv), st))

    def counts(self):
        frames = np.zeros(self.nchan, np.int32)
        consumed = np.zeros(self.nchan, np.int64)
        _ck(self.L.b200fsk_counts(self.h, frames.ctypes.data, consumed.ctypes.data))
        return frames, consumed

    # -- statistics ----------------------------------------------------------------------------------------
    def stats(self, chan: int = 0) -> ModemStats:
        s = _Stats()
        _ck(self.L.b200fsk_get_stats(self.h, chan, C.byref(s)))
        eye = np.ctypeslib.as_array(s.rx_eye)[: s.neyetr, : s.neyesamp].copy()
        return ModemStats(s.snr_est, s.EbNodB, s.ppm, s.rx_timing, s.norm_rx_timing, s.foff, list(s.f_est)[: self.M],
                          s.nin, eye, s.frames, s.samples)

    def Sf(self, chan: int = 0) -> np.ndarray:
        out = np.zeros(self.Ndft, np.float32)
        _ck(self.L.b200fsk_get_Sf(self.h, chan, out.ctypes.data))
        return out

    @property
    def launches(self) -> int:
        return int(self.L.b200fsk_launch_count(self.h))

    def last_kernel_ms(self) -> float:
        return float(self.L.b200fsk_last_kernel_ms(self.h))

    KERNEL_NAMES = {-1: "none", 0: "fsk_demod_kernel", 1: "fsk_demod_pipe_kernel", 2: "fsk_wpc_kernel", 3: "fsk_duo_kernel"}

    def last_kernel(self) -> str:
        """name of the frame kernel the most recent launch used (include/b200fsk.h: b200fsk_last_kernel_kind)"""
        return self.KERNEL_NAMES[int(self.L.b200fsk_last_kernel_kind(self.h))]


# ==================================================================================================
# Single-channel functions with the reference's names (utils/fsk.h:105-205)
# ==================================================================================================
class FSK:
    """Stand-in for ``struct FSK *``.  The GPU handle is created lazily on the first demod call so that
    ``fsk_set_freq_est_limits`` / ``fsk_set_freq_est_alg`` can be applied after create, as in the CLI."""

    def __init__(self, Fs, Rs, M, P, Nsym, f1_tx, fs_tx, device=0):
        if not (Fs > 0 and Rs > 0 and f1_tx > 0 and fs_tx > 0 and P > 0 and Nsym > 0):
            raise AssertionError("fsk_create: non-positive parameter")          # fsk.c:120-125
        if Fs % Rs or (Fs // Rs) % P or M not in (2, 4):
            raise AssertionError("fsk_create: Fs%Rs, (Fs/Rs)%P, M in {2,4}")       # fsk.c:127-130
        self.Fs, self.Rs, self.mode, self.P, self.Nsym, self.f1_tx, self.fs_tx = Fs, Rs, M, P, Nsym, f1_tx, fs_tx
        self.Ts = Fs // Rs
        self.N = self.Ts * Nsym
        self.Nmem = self.N + 2 * self.Ts
        self.Nbits = Nsym if M == 2 else 2 * Nsym
        self.est_min, self.est_max, self.freq_est_type = 0, Fs, 0                # fsk.c:156-157
        self._device = device
        self._b: FskBatch | None = None
        self._nin = self.N

    def _batch(self) -> FskBatch:
        if self._b is None:
            lo = max(self.est_min, -self.Fs // 2)
            hi = min(self.est_max, self.Fs // 2)
            self._b = FskBatch(self.Fs, self.Rs, self.mode, self.P, self.Nsym, self.f1_tx, self.fs_tx, lo, hi,
                               self.freq_est_type, 1, self._device)
            self._b.enable_stats(True)
            self.Ndft = self._b.Ndft
        return self._b


def fsk_create(Fs, Rs, M, tx_f1, tx_fs) -> FSK:                                   # fsk.c:215
    return FSK(Fs, Rs, M, FSK_DEFAULT_P, FSK_DEFAULT_NSYM, tx_f1, tx_fs)


def fsk_create_hbr(Fs, Rs, M, P, Nsym, tx_f1, tx_fs) -> FSK:                      # fsk.c:235
    return FSK(Fs, Rs, M, P, Nsym, tx_f1, tx_fs)


def fsk_destroy(fsk: FSK) -> None:                                                # fsk.c:249
    if fsk._b is not None:
        fsk._b.close()
        fsk._b = None


def fsk_set_freq_est_limits(fsk: FSK, est_min: int, est_max: int) -> None:        # fsk.c:1017
    assert est_min >= -fsk.Fs // 2 and est_max <= fsk.Fs // 2 and est_max > est_min
    assert fsk._b is None, "limits must be set before the first demod call"
    fsk.est_min, fsk.est_max = est_min, est_max


def fsk_set_freq_est_alg(fsk: FSK, est_type: int) -> None:                        # fsk.c:1031
    assert fsk._b is None, "estimator must be chosen before the first demod call"
    fsk.freq_est_type = int(bool(est_type))


def fsk_nin(fsk: FSK) -> int:                                                     # fsk.c:424
    return fsk._nin


def _demod(fsk: FSK, fsk_in: np.ndarray, soft: bool) -> np.ndarray:
    b = fsk._batch()
    x = np.ascontiguousarray(fsk_in)
    if np.iscomplexobj(x):
        x = x.astype(np.complex64).view(np.float32)
    x = x.astype(np.float32, copy=False).reshape(-1)
    if x.size != 2 * fsk._nin:
        raise ValueError(f"fsk_demod needs exactly fsk_nin()={fsk._nin} samples")
    out, frames = b.push(x.reshape(1, -1), "f32", soft=soft, max_frames=1)
    assert frames[0] == 1
    fsk._nin = int(b.nin()[0])
    return out[0, : fsk.Nbits].copy()


def fsk_demod(fsk: FSK, fsk_in: np.ndarray) -> np.ndarray:                        # fsk.c:930
    """COMP fsk_in[nin] -> uint8 rx_bits[Nbits]."""
    return _demod(fsk, fsk_in, False)


def fsk_demod_sd(fsk: FSK, fsk_in: np.ndarray) -> np.ndarray:                     # fsk.c:934
    """COMP fsk_in[nin] -> float rx_sd[Nbits]."""
    return _demod(fsk, fsk_in, True)


def fsk_get_demod_stats(fsk: FSK) -> ModemStats:                                  # fsk.c:991
    return fsk._batch().stats(0)


def fsk_clear_estimators(fsk: FSK) -> None:                                       # fsk.c:981
    fsk._batch().clear_estimators()
    fsk._nin = fsk.N


# ==================================================================================================
# Mixed batches (BASELINE.json configs[3]): channels with different modem parameters
# ==================================================================================================
class FskMixedBatch:
    """A set of channels with assorted (Fs, Rs, M, P, Nsym, limits, estimator) settings.  Channels that share a
    configuration are grouped into one :class:`FskBatch` handle (one kernel launch per group and push); channel
    order is preserved in the results.  `configs[c]` are keyword arguments of :class:`FskBatch` (without nchan)."""

    def __init__(self, configs: list[dict], device: int = 0):
        self.configs = [dict(c) for c in configs]
        self.groups: dict[tuple, list[int]] = {}
        for c, cfg in enumerate(self.configs):
            self.groups.setdefault(tuple(sorted(cfg.items())), []).append(c)
        self.batches = {k: FskBatch(**dict(k), nchan=len(ch), device=device) for k, ch in self.groups.items()}

    def push(self, data: list[np.ndarray], fmt: str = "cs16", soft: bool = True, softinv: bool = False):
        """data[c]: the new samples of channel c (channels of one group must bring equally many).  Returns per channel
        the 1-D array of bits / soft decisions produced by this push."""
        res: list = [None] * len(self.configs)
        for k, ch in self.groups.items():
            b = self.batches[k]
            block = np.stack([np.ascontiguousarray(data[c]).reshape(-1) for c in ch])
            out, frames = b.push(block, fmt, soft=soft, softinv=softinv)
            for i, c in enumerate(ch):
                res[c] = out[i, : frames[i] * b.Nbits].copy()
        return res

    def close(self):
        for b in self.batches.values():
            b.close()


# ==================================================================================================
# Soft-bit header sync + RS41 frame assembly on the device
# ==================================================================================================
class _Rs41Frame(C.Structure):
    _fields_ = [("chan", C.c_int), ("len", C.c_int), ("nbytes_read", C.c_int), ("first_bit", C.c_longlong), ("score", C.c_float),
                ("bytes", C.c_ubyte * 518), ("bytescore", C.c_float * 518), ("ec", C.c_int), ("crc_mask", C.c_int)]


def rs41_framesync_device(d_soft, stride: int, nbits, nchan: int, softinv: bool = False, inv: bool = True, device: int = 0,
                          stream=None, ecc: int = 0):
    """What `rs41mod --softin [-i] -r` extracts from fsk_demod's soft decisions, run on soft bits that are still in HBM.
    d_soft: torch CUDA float32 tensor [nchan, stride] (or device address); nbits: int (all channels) or torch int32 tensor.
    ecc: 0 = frames before the Reed-Solomon stage, 1 = `--ecc` (rs41mod's default), 2 = `--ecc2` (what `--json` selects).
    Returns a list of (chan, len, first_bit, score, bytes[len] (numpy uint8), nbytes_read, ec, crc_mask)."""
    L = lib()
    L.b200fsk_rs41_frames_device.argtypes = [C.c_void_p, C.c_long, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                             C.c_int, C.POINTER(C.c_int), C.c_int, C.c_void_p]
    L.b200fsk_framesync_last_error.restype = C.c_char_p
    ptr = d_soft.data_ptr() if hasattr(d_soft, "data_ptr") else int(d_soft)
    if hasattr(nbits, "data_ptr"):
        pn, nmax = C.c_void_p(nbits.data_ptr()), int(nbits.max().item())
    else:
        pn, nmax = None, int(nbits)
    cap = nchan * (nmax // 4144 + 2)
    fr = (_Rs41Frame * cap)()
    n = C.c_int(0)
    st = C.c_void_p(stream.cuda_stream) if stream is not None and hasattr(stream, "cuda_stream") else None
    rc = L.b200fsk_rs41_frames_device(C.c_void_p(ptr), stride, pn, nmax, nchan, int(softinv), int(inv), int(ecc), fr, cap, C.byref(n), device, st)
    if rc:
        raise B200Error(f"rs41 framesync error {rc}: {L.b200fsk_framesync_last_error().decode()}")
    return [(f.chan, f.len, f.first_bit, f.score, np.frombuffer(bytes(f.bytes), np.uint8)[: f.len].copy(), f.nbytes_read, f.ec, f.crc_mask) for f in fr[: n.value]]


class Rs41FrameSync:
    """Streaming `rs41mod --softin [-i] [--ecc2]` front for ``nchan`` channels: push() soft bits as fsk_demod produces them (device
    memory), frames come back from the push that completes them; ``flush=True`` on the last push applies the reference's EOF rule."""

    def __init__(self, nchan: int, softinv: bool = False, inv: bool = True, ecc: int = 0, device: int = 0):
        L = lib()
        L.b200fsk_rs41sync_create.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        L.b200fsk_rs41sync_push_device.argtypes = [C.c_void_p, C.c_void_p, C.c_long, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int,
                                                   C.POINTER(C.c_int), C.c_void_p]
        L.b200fsk_rs41sync_destroy.argtypes = [C.c_void_p]
        L.b200fsk_rs41sync_destroy.restype = None
        L.b200fsk_framesync_last_error.restype = C.c_char_p
        self.L, self.nchan, self.h = L, nchan, C.c_void_p()
        rc = L.b200fsk_rs41sync_create(C.byref(self.h), nchan, int(softinv), int(inv), int(ecc), device)
        if rc:
            raise B200Error(f"rs41sync_create error {rc}: {L.b200fsk_framesync_last_error().decode()}")

    def push(self, d_soft, stride: int, nbits, flush: bool = False, stream=None):
        """-> list of (chan, len, first_bit, score, bytes[len], nbytes_read, ec, crc_mask) completed by this push"""
        if hasattr(nbits, "data_ptr"):
            pn, nmax = C.c_void_p(nbits.data_ptr()), int(nbits.max().item())
        else:
            pn, nmax = None, int(nbits)
        cap = self.nchan * (nmax // 4144 + 3)
        fr = (_Rs41Frame * cap)()
        n = C.c_int(0)
        ptr = d_soft.data_ptr() if hasattr(d_soft, "data_ptr") else int(d_soft)
        st = C.c_void_p(stream.cuda_stream) if stream is not None and hasattr(stream, "cuda_stream") else None
        rc = self.L.b200fsk_rs41sync_push_device(self.h, C.c_void_p(ptr), stride, pn, nmax, int(flush), fr, cap, C.byref(n), st)
        if rc:
            raise B200Error(f"rs41sync_push error {rc}: {self.L.b200fsk_framesync_last_error().decode()}")
        return [(f.chan, f.len, f.first_bit, f.score, np.frombuffer(bytes(f.bytes), np.uint8)[: f.len].copy(), f.nbytes_read, f.ec, f.crc_mask)
                for f in fr[: n.value]]

    def close(self):
        if getattr(self, "h", None):
            self.L.b200fsk_rs41sync_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def rs41_framesync_kernel_ms() -> float:
    """Device time of the frame-sync (+ECC) kernel in the most recent rs41_framesync_device call of this thread."""
    L = lib()
    L.b200fsk_framesync_last_kernel_ms.restype = C.c_float
    return float(L.b200fsk_framesync_last_kernel_ms())
