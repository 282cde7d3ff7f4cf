"""The C-ABI libraries load and export every function declared in include/*.h (no compute calls: runs without a GPU);
creating a handle without a CUDA device must fail loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared(header):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(b200(?:fsk|dft)_\w+)\s*\(", txt)))


@pytest.mark.parametrize("header,libname", [("b200fsk.h", "libb200fsk.so"), ("b200dft.h", "libb200dft.so")])
def test_every_declared_symbol_is_exported(built, header, libname):
    lib = C.CDLL(os.path.join(ROOT, "radiosonde_auto_rx_b200", libname))
    names = declared(header)
    assert len(names) >= 8
    for n in names:
        assert hasattr(lib, n), f"{libname} does not export {n}"


def test_create_without_gpu_fails_loudly(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from radiosonde_auto_rx_b200 import dft, fsk
    with pytest.raises(fsk.B200Error):
        fsk.FskBatch(48000, 4800)
    with pytest.raises(dft.B200DftError):
        dft.DftDetectBatch(48000, 15.0)


def test_bad_configuration_is_an_error_code_not_a_crash(built):
    """The reference asserts on these (fsk.c:120-130); the C-ABI returns B200FSK_EINVAL before touching CUDA."""
    from radiosonde_auto_rx_b200 import fsk
    for kw in (dict(Fs=48000, Rs=4700), dict(Fs=48000, Rs=4800, M=3), dict(Fs=48000, Rs=4800, P=7)):
        with pytest.raises(fsk.B200Error, match="error -1"):
            fsk.FskBatch(**kw)


def test_cli_binaries_fail_loudly_without_gpu(built):
    """the drop-in executables have no CPU path either: without a CUDA device they say so and exit non-zero (no output bytes)"""
    import subprocess
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    bin_dir = os.path.join(ROOT, "radiosonde_auto_rx_b200", "bin")
    for args, msg in ((["fsk_demod", "--cs16", "2", "48000", "4800", "-", "-"], b"no CPU fallback"),
                      (["dft_detect", "--iq", "--bw", "15", "--dc", "-", "48000", "16"], b"no CPU fallback"),
                      (["dft_detect", "--IQ", "0.1", "--bw", "15", "-", "2400000", "16"], b"no CPU fallback")):
        r = subprocess.run([os.path.join(bin_dir, args[0])] + args[1:], input=bytes(4096), capture_output=True, timeout=60)
        assert r.returncode != 0 and r.stdout == b"" and msg in r.stderr, (args, r.returncode, r.stderr[-200:])
