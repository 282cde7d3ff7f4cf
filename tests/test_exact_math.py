"""csrc/exact_math.h: em_atan2f (the rounding-explicit fdlibm atan2f the GPU evaluates once per frame, utils/fsk.c:707) against the
system libm the reference links, bit for bit: 1.2e7 seeded arguments over every exponent range, plus signed zeros, infinities, NaN,
subnormals and the axis / diagonal cases.  The GPU side (device == host) is tests/test_fsk_gpu.py::test_exact_math_device_equals_libm."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def shim(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("em") / "libem_shim.so")
    r = subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-o", so, os.path.join(ROOT, "tests", "em_shim.c"), "-lm"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    L = C.CDLL(so)
    for f in (L.em_atan2f_vec, L.libm_atan2f_vec):
        f.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_long]
    return L


def both(L, y, x):
    y, x = np.ascontiguousarray(y, np.float32), np.ascontiguousarray(x, np.float32)
    a, b = np.empty_like(y), np.empty_like(y)
    L.em_atan2f_vec(y.ctypes.data, x.ctypes.data, a.ctypes.data, y.size)
    L.libm_atan2f_vec(y.ctypes.data, x.ctypes.data, b.ctypes.data, y.size)
    return a.view(np.uint32), b.view(np.uint32)


def same(a, b, y, x):
    nan_a, nan_b = np.isnan(a.view(np.float32)), np.isnan(b.view(np.float32))
    bad = (a != b) & ~(nan_a & nan_b)                    # NaN payloads may differ, NaN-ness may not
    assert not bad.any(), (int(bad.sum()), y[bad][:4], x[bad][:4], a[bad][:4], b[bad][:4])


def test_random_arguments_all_magnitudes(shim):
    rng = np.random.default_rng(20260923)
    n = 4_000_000
    # (1) the operating range of the timing estimator: both components O(1e-3 .. 1e3)
    y = rng.standard_normal(n).astype(np.float32) * np.float32(10) ** rng.uniform(-3, 3, n).astype(np.float32)
    x = rng.standard_normal(n).astype(np.float32) * np.float32(10) ** rng.uniform(-3, 3, n).astype(np.float32)
    same(*both(shim, y, x), y, x)
    # (2) raw bit patterns: every exponent incl. subnormals, infinities and NaNs
    y = rng.integers(0, 2 ** 32, n, dtype=np.uint64).astype(np.uint32).view(np.float32)
    x = rng.integers(0, 2 ** 32, n, dtype=np.uint64).astype(np.uint32).view(np.float32)
    same(*both(shim, y, x), y, x)
    # (3) ratios near the breakpoints of the argument reduction (7/16, 11/16, 19/16, 39/16) and huge / tiny quotients
    x = (rng.uniform(0.5, 2.0, n)).astype(np.float32) * rng.choice(np.float32([-1, 1]), n)
    r = rng.choice(np.float32([7 / 16, 11 / 16, 19 / 16, 39 / 16, 1.0, 2.0 ** -26, 2.0 ** 26, 2.0 ** 60, 2.0 ** -60]), n)
    y = (x * r * (1 + rng.uniform(-1e-6, 1e-6, n).astype(np.float32))).astype(np.float32) * rng.choice(np.float32([-1, 1]), n)
    same(*both(shim, y, x), y, x)


def test_special_values(shim):
    sp = np.float32([0.0, -0.0, 1.0, -1.0, np.inf, -np.inf, np.nan, 1e-45, -1e-45, 1.17549435e-38, 3.4028235e38, -3.4028235e38,
                     0.5, 2.0, 1e-20, 1e20, 0.4375, 0.6875, 1.1875, 2.4375])
    y, x = np.meshgrid(sp, sp)
    same(*both(shim, y.ravel(), x.ravel()), y.ravel(), x.ravel())
