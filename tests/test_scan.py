"""scanner peak stage.  CPU: the C detect_peaks / candidate list against outputs of the reference's own
autorx.utils.detect_peaks (tests/golden/golden_peaks.json, generator: tests/golden/make_golden_peaks.py) and the numpy restatement
of scan.py:1005-1063.  GPU: periodogram kernel against float64 numpy, and capture -> spectrum -> candidates -> `--IQ` detector."""
import importlib.util
import json
import os

import numpy as np
import pytest

from oracle import pyoracle as po
from radiosonde_auto_rx_b200 import dft, siggen

HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location("make_golden_peaks", os.path.join(HERE, "golden", "make_golden_peaks.py"))
mgp = importlib.util.module_from_spec(spec)
spec.loader.exec_module(mgp)
GP = json.load(open(os.path.join(HERE, "golden", "golden_peaks.json")))


@pytest.mark.parametrize("case", mgp.CASES, ids=lambda c: f"seed{c['seed']}")
def test_detect_peaks_matches_reference_function(built, case):
    x = mgp.spectrum(case["seed"], case["n"], case["npeaks"], case["flat"], case["nans"])
    g = GP[str(case["seed"])]
    assert abs(float(np.nansum(x)) - g["xsum"]) < 1e-6                       # same input as when the golden file was made
    assert list(po.detect_peaks_oracle(x, g["mph"], g["mpd"])) == g["ind"]
    assert list(dft.detect_peaks(x, g["mph"], g["mpd"])) == g["ind"]


@pytest.mark.parametrize("seed", range(20, 32))
def test_candidate_list_matches_oracle(built, seed):
    rng = np.random.default_rng(seed)
    n, step = 2500, 800.0
    power = mgp.spectrum(seed, n, int(rng.integers(0, 25)))
    freq = 400.05e6 + np.arange(n) * step
    kw = dict(snr_threshold=float(rng.choice([6.0, 10.0])), min_distance=float(rng.choice([1000.0, 5000.0])), quantization=10000.0,
              min_freq_hz=400.3e6, max_freq_hz=401.7e6, max_peaks=int(rng.choice([3, 10, 100])))
    exp, nf = po.scan_candidates_oracle(freq, power, step, **kw)
    got, gnf = dft.scan_candidates(freq, power, step, **kw)
    assert gnf == nf and list(got) == list(exp)


@pytest.mark.gpu
def test_psd_kernel_against_numpy(cuda):
    import torch
    sr = 2400000
    wide = siggen.wideband_mix(31, sr // 2, sr, [("RS41", 300e3), ("DFM9", -500e3), ("RS92", 850e3)], snr_db=8.0, dc=(0.0, 0.0))
    d = torch.from_numpy(wide).cuda()
    for nfft in (1024, 4096, 16384):
        f, p, step, ms = dft.scan_psd(d, "cs16", wide.shape[0], sr, 402.0e6, nfft)
        ref = po.psd_oracle(wide, nfft)
        assert step == sr / nfft and f[0] == 402.0e6 - sr / 2 and abs(f[nfft // 2] - 402.0e6) < 1e-6
        assert np.abs(p - ref).max() < 2e-3, (nfft, np.abs(p - ref).max())    # float32 FFT vs float64, in dB


@pytest.mark.gpu
def test_scan_then_detect_on_one_capture(cuda):
    """scan.py's flow on one buffer: spectrum -> candidates (strongest first) -> one `dft_detect --IQ` per candidate, all on the GPU"""
    import torch
    sr, fc = 2400000, 403.0e6
    emit = [("RS41", 300e3), ("DFM9", -500e3), ("RS92", -900e3), ("LMS6", 1000e3)]
    wide = siggen.wideband_mix(33, int(sr * 1.2), sr, emit, snr_db=6.0, dc=(100.0, -80.0))
    d = torch.from_numpy(wide).cuda()
    f, p, step, _ = dft.scan_psd(d, "cs16", wide.shape[0], sr, fc, 4096)
    exp, _ = po.scan_candidates_oracle(f, po.psd_oracle(wide, 4096), step, max_peaks=8)
    cand, nf = dft.scan_candidates(f, p, step, max_peaks=8)
    assert list(cand) == list(exp)
    found = {round((c - fc) / 1e4) * 1e4 for c in cand}
    assert {e[1] for e in emit} <= found, (found, cand)
    b = dft.DftWidebandBatch(sr, 15.0, True, nchan=len(cand))
    res = b.run_device(d, "cs16", wide.shape[0], (cand - fc) / sr, cont=False)
    types = {round((c - fc) / 1e4) * 1e4: (r[1][0].type if r[1] else None) for c, r in zip(cand, res)}
    for kind, hz in emit:
        assert types[hz] is not None and types[hz][:3] == kind[:3], (kind, hz, types)
        rc, dets, _, _ = po.dft_oracle_run_wideband(wide, sr, hz / sr, 15.0, True, False)
        k = list(cand).index(fc + hz)
        assert res[k][0] == rc and [(x.type, x.pos) for x in res[k][1]] == [(o[0], o[2]) for o in dets]
