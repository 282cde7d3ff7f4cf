#!/usr/bin/env python
"""Golden vectors for the scanner peak stage: runs the UNMODIFIED reference function
autorx.utils.detect_peaks (auto_rx/autorx/utils.py:435) imported from /root/reference (modules the reference imports but this image
lacks are stubbed -- they are not used by detect_peaks) on seeded spectra, and records inputs' seeds + outputs in golden_peaks.json.
    python tests/golden/make_golden_peaks.py"""
import json, os, sys, types
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def spectrum(seed: int, n: int, npeaks: int, flat: bool = False, nans: int = 0) -> np.ndarray:
    """dB-like spectrum: noise floor + a few Gaussian-shaped carriers; optional flat-topped peaks and NaNs (rtl_power drop-outs)."""
    rng = np.random.default_rng(seed)
    x = -70.0 + rng.standard_normal(n) * 1.5
    for _ in range(npeaks):
        c, h, w = int(rng.integers(5, n - 5)), float(rng.uniform(6, 40)), float(rng.uniform(1.0, 6.0))
        x += h * np.exp(-0.5 * ((np.arange(n) - c) / w) ** 2)
    x = np.round(x, 2)                                        # rtl_power prints two decimals: exact ties do occur
    if flat:
        for c in rng.integers(5, n - 5, 4):
            x[c:c + 3] = x[c:c + 3].max()
    for c in rng.integers(1, n - 1, nans):
        x[c] = np.nan
    return x


CASES = [dict(seed=s, n=n, npeaks=k, flat=f, nans=q, mph_over=m, mpd=d)
         for s, n, k, f, q, m, d in [(1, 600, 6, False, 0, 10.0, 1.25), (2, 600, 12, False, 0, 10.0, 1.25), (3, 2048, 30, False, 0, 8.0, 4.0),
                                     (4, 300, 5, True, 0, 10.0, 1.25), (5, 300, 8, True, 3, 6.0, 3.0), (6, 64, 2, False, 0, None, 1.0),
                                     (7, 1000, 40, True, 5, 3.0, 12.5), (8, 3, 0, False, 0, None, 1.0), (9, 2, 0, False, 0, None, 1.0),
                                     (10, 12000, 60, False, 0, 10.0, 1.25)]]


def main():
    for m in ("semver", "requests", "dateutil", "dateutil.parser"):
        if m not in sys.modules:
            try:
                __import__(m)
            except Exception:
                sys.modules[m] = types.ModuleType(m)
    sys.path.insert(0, "/root/reference/auto_rx")
    from autorx.utils import detect_peaks
    out = {}
    for c in CASES:
        x = spectrum(c["seed"], c["n"], c["npeaks"], c["flat"], c["nans"])
        mph = None if c["mph_over"] is None else float(np.nanmedian(x) + c["mph_over"])
        ind = detect_peaks(x, mph=mph, mpd=c["mpd"], show=False)
        out[str(c["seed"])] = {"mph": mph, "mpd": c["mpd"], "ind": [int(i) for i in ind], "xsum": float(np.nansum(x))}
    json.dump(out, open(os.path.join(HERE, "golden_peaks.json"), "w"), indent=1)
    print("wrote golden_peaks.json", {k: len(v["ind"]) for k, v in out.items()})


if __name__ == "__main__":
    main()
