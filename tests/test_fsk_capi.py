"""The reference's own single-channel API (utils/fsk.h:105-205) exported unchanged by libb200fsk.so (include/fsk.h).

* CPU: every fsk.h function is an exported symbol; the modulators (host arithmetic) are bit-identical to the reference's;
  the UNMODIFIED reference program utils/fsk_demod.c builds against include/ + -lb200fsk (oracle/Makefile: _ref/fsk_demod_capi)
  and, having no CPU demodulator behind it, fails loudly without a CUDA device.
* GPU: that binary -- the reference's main(), our library underneath -- produces the same stdout bytes (and the same statistics
  text) as the all-reference binary oracle/_ref/fsk_demod on the same stdin."""
import ctypes as C
import json
import os
import re
import subprocess

import numpy as np
import pytest

from oracle import pyoracle as po
from radiosonde_auto_rx_b200 import siggen

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "radiosonde_auto_rx_b200", "libb200fsk.so")
CAPI = os.path.join(po.REF_DIR, "fsk_demod_capi")
REFBIN = os.path.join(po.REF_DIR, "fsk_demod")
needs_capi = pytest.mark.skipif(not (os.path.exists(CAPI) and os.path.exists(REFBIN)), reason="oracle/_ref/fsk_demod_capi absent")


def fsk_h_functions():
    txt = open(os.path.join(ROOT, "include", "fsk.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(fsk_\w+)\s*\(", txt)))


def test_fsk_h_symbols_exported(built):
    lib = C.CDLL(LIB)
    names = fsk_h_functions()
    assert set(names) >= {"fsk_create", "fsk_create_hbr", "fsk_destroy", "fsk_nin", "fsk_demod", "fsk_demod_sd", "fsk_set_freq_est_limits",
                          "fsk_set_freq_est_alg", "fsk_clear_estimators", "fsk_get_demod_stats", "fsk_stats_normalise_eye",
                          "fsk_enable_burst_mode", "fsk_mod", "fsk_mod_c", "fsk_mod_ext_vco"}
    for n in names:
        assert hasattr(lib, n), n


@pytest.mark.skipif(not os.path.exists(os.path.join(po.REF_DIR, "libfsk_ref.so")), reason="oracle/_ref absent")
@pytest.mark.parametrize("M", [2, 4])
def test_modulators_equal_the_reference(built, M):
    """fsk_mod_c / fsk_mod / fsk_mod_ext_vco over several frames (the TX phase carries over): bit-identical to utils/fsk.c:266-411"""
    ours, ref = C.CDLL(LIB), C.CDLL(os.path.join(po.REF_DIR, "libfsk_ref.so"))
    for L in (ours, ref):
        L.fsk_create_hbr.restype = C.c_void_p
        L.fsk_create_hbr.argtypes = [C.c_int] * 7
        for f in ("fsk_mod_c", "fsk_mod", "fsk_mod_ext_vco"):
            getattr(L, f).argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.fsk_destroy.argtypes = [C.c_void_p]
    Fs, Rs, P, Nsym = 48000, 4800, 5, 64
    a, b = ours.fsk_create_hbr(Fs, Rs, M, P, Nsym, 1200, 2400), ref.fsk_create_hbr(Fs, Rs, M, P, Nsym, 1200, 2400)
    rng = np.random.default_rng(3)
    N = Nsym * (Fs // Rs)
    for _ in range(5):
        bits = rng.integers(0, 2, Nsym * (1 if M == 2 else 2)).astype(np.uint8)
        for f, shape in (("fsk_mod_c", (N, 2)), ("fsk_mod", (N,)), ("fsk_mod_ext_vco", (N,))):
            oa, ob = np.zeros(shape, np.float32), np.zeros(shape, np.float32)
            getattr(ours, f)(a, oa.ctypes.data, bits.ctypes.data)
            getattr(ref, f)(b, ob.ctypes.data, bits.ctypes.data)
            assert oa.tobytes() == ob.tobytes(), f
    ours.fsk_destroy(a); ref.fsk_destroy(b)


@pytest.mark.skipif(not os.path.isdir("/root/reference/utils"), reason="reference tree absent (GPU box)")
def test_unmodified_reference_main_builds_against_our_headers(built):
    """oracle/Makefile compiles /root/reference/utils/fsk_demod.c with -I include -lb200fsk; the reference file is untouched"""
    r = subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "-B", "_ref/fsk_demod_capi"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert os.path.exists(CAPI)
    needed = subprocess.run(["ldd", CAPI], capture_output=True, text=True).stdout
    assert "libb200fsk.so" in needed and "libfsk_ref" not in needed
    # the headers it was compiled against are OURS (the file goes through stdin: a quoted include would otherwise find the
    # reference's fsk.h next to it)
    deps = subprocess.run("gcc -M -w -I" + os.path.join(ROOT, "include") + " -x c - < /root/reference/utils/fsk_demod.c", shell=True,
                          capture_output=True, text=True, cwd=os.path.join(ROOT, "oracle")).stdout
    assert os.path.join(ROOT, "include", "fsk.h") in deps.replace("../include", os.path.join(ROOT, "include")) or "include/fsk.h" in deps
    assert "/root/reference/utils/fsk.h" not in deps and "/root/reference/utils/modem_stats.h" not in deps


@needs_capi
def test_capi_binary_fails_loudly_without_gpu(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    iq = siggen.fsk_iq(1, 20_000)
    r = subprocess.run([CAPI, "--cs16", "2", "48000", "4800", "-", "-"], input=iq.tobytes(), capture_output=True, timeout=60)
    assert r.returncode != 0 and r.stdout == b"" and b"no CPU fallback" in r.stderr


@needs_capi
@pytest.mark.gpu
@pytest.mark.parametrize("args", [
    ["--cs16", "-b", "-5000", "-u", "5000", "-s", "--mask", "5000", "--nsym=300", "-p", "5", "--stats=5", "2", "48000", "4800", "-", "-"],   # RS41, decode.py:901
    ["--cs16", "-b", "-20000", "-u", "20000", "-s", "--stats=5", "2", "48000", "4800", "-", "-"],                                           # RS92
    ["--cs16", "2", "48000", "4800", "-", "-"],                                                                                               # defaults, hard bits
    ["--cs16", "-s", "-i", "-t", "4", "48000", "4800", "-", "-"],                                                                             # 4FSK, soft inverted
])
def test_reference_main_over_our_library_equals_reference_binary(cuda, args):
    M = 4 if "4" in args[-5:-4] else 2
    iq = siggen.fsk_iq(23, 120_000, M=M, ebno_db=11.0, ppm=-30.0, spacing=4800.0 if M == 4 else None)
    ref = subprocess.run([REFBIN] + args, input=iq.tobytes(), capture_output=True, timeout=300)
    got = subprocess.run([CAPI] + args, input=iq.tobytes(), capture_output=True, timeout=300)
    assert ref.returncode == got.returncode == 0, got.stderr[-400:]
    assert got.stdout == ref.stdout and len(ref.stdout) > 0, "stdout of the reference main over libb200fsk differs from the reference binary"

    def parse(b):
        out = []
        for l in b.decode().splitlines()[1:]:
            d = json.loads(l.replace("-nan", "NaN").replace("nan", "NaN"))
            d.pop("eye_diagram", None)          # the reference indexes f_int at -1 on some frames (fsk.c:884-887): stack garbage
            out.append(d)
        return out
    assert got.stderr.splitlines()[0] == ref.stderr.splitlines()[0]
    assert json.dumps(parse(got.stderr)) == json.dumps(parse(ref.stderr)), "statistics printed by the reference main differ"
