"""Oracle vs the committed golden vectors (outputs of the UNMODIFIED reference, tests/golden/make_golden.py).
Runs anywhere (no GPU, no /root/reference): this is what pins the oracle on the GPU box."""
import hashlib
import json
import os
import sys

import numpy as np
import pytest

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
sys.path.insert(0, HERE)
import make_golden as mg                               # noqa: E402
from oracle import pyoracle as po                      # noqa: E402

G = np.load(os.path.join(HERE, "golden_fsk.npz"))
GD = json.load(open(os.path.join(HERE, "golden_dft.json")))
GW = json.load(open(os.path.join(HERE, "golden_dft_wb.json")))
G96 = json.load(open(os.path.join(HERE, "golden_dft96.json")))


@pytest.mark.parametrize("name", list(mg.FSK_CASES))
def test_fsk_oracle_reproduces_golden(built, name):
    cfg = po.FskConfig(**mg.FSK_CASES[name][0])
    iq = mg.fsk_input(name)
    assert hashlib.sha256(iq.tobytes()).digest() == G[name + "/sha"].tobytes(), "synthetic input drifted from the fixture"
    out, _, recs = po.Oracle(cfg).run_stream(iq)
    assert out.tobytes() == G[name + "/out"].tobytes()
    assert np.array_equal(recs["nin_next"], G[name + "/nin"])
    assert np.array_equal(recs["ppm"].view(np.uint32), G[name + "/ppm"].view(np.uint32))
    assert np.array_equal(recs["snr_est"].view(np.uint32), G[name + "/snr_est"].view(np.uint32))
    assert np.array_equal(recs["f_est"].view(np.uint32), G[name + "/f_est"].view(np.uint32))


@pytest.mark.parametrize("name", list(mg.DFT_CASES))
def test_dft_oracle_reproduces_golden(built, name):
    iq = mg.dft_input(name)
    assert hashlib.sha256(iq.tobytes()).hexdigest() == GD[name]["sha"]
    rc, dets, _ = po.dft_oracle_run(iq, 48000, 15.0, True, True)
    assert rc == GD[name]["exit_code"] and len(dets) == len(GD[name]["dets"])
    for a, b in zip(dets, GD[name]["dets"]):
        assert a[0] == b[0] and a[2] == b[2] and abs(a[1] - b[1]) < 1e-3 and abs(a[3] - b[3]) < 0.2


@pytest.mark.parametrize("name", list(mg.DFT96_CASES))
def test_dft_96khz_oracle_reproduces_golden(built, name):
    """the 96 kHz IF of auto_rx's wideband sondes (scan.py:476-482): N_DFT 16384, 97/193 taps"""
    iq = mg.dft96_input(name)
    g = G96[name]
    assert hashlib.sha256(iq.tobytes()).hexdigest() == g["sha"]
    rc, dets, _ = po.dft_oracle_run(iq, 96000, g["bw"], True, True)
    assert rc == g["exit_code"] and len(dets) == len(g["dets"])
    for a, b in zip(dets, g["dets"]):
        assert a[0] == b[0] and a[2] == b[2] and abs(a[1] - b[1]) < 1e-3 and abs(a[3] - b[3]) < 0.2


@pytest.mark.gpu
def test_dft_96khz_cuda_reproduces_golden(cuda):
    """`dft_detect -c -v --iq --bw {64|32} --dc - 96000 16` on the GPU against the reference binary's recorded output"""
    from radiosonde_auto_rx_b200 import dft
    ndet = 0
    for bw in sorted({g["bw"] for g in G96.values()}):
        names = [n for n in mg.DFT96_CASES if G96[n]["bw"] == bw]
        iq = np.stack([mg.dft96_input(n) for n in names])
        b = dft.DftDetectBatch(96000, bw, True, nchan=len(names))
        assert b.N_DFT == 16384
        got = b.run(iq, cont=True)
        for (rc, dets), n in zip(got, names):
            g = G96[n]
            assert rc == g["exit_code"] and len(dets) == len(g["dets"]), (n, rc, dets, g)
            for d, e in zip(dets, g["dets"]):
                assert d.type == e[0] and d.pos == e[2] and abs(d.score - e[1]) < 2e-4 and abs(d.df_hz - e[3]) < 0.2
            ndet += len(dets)
    assert ndet >= 8


@pytest.mark.parametrize("name", list(mg.DFT_WB_CASES))
def test_dft_wideband_oracle_reproduces_golden(built, name):
    """`dft_detect --IQ <fq>` (translate + decimate front) restatement against the reference binary's recorded output"""
    iq = mg.dft_wb_input(name)
    assert hashlib.sha256(iq.tobytes()).hexdigest() == GW[name]["sha"]
    sr = mg.DFT_WB_CASES[name][0]
    ndet = 0
    for f, g in GW[name]["fq"].items():
        rc, dets, _, _ = po.dft_oracle_run_wideband(iq, sr, float(f) / sr, 15.0, True, True)
        assert rc == g["exit_code"] and len(dets) == len(g["dets"]), (f, rc, dets, g)
        for a, b in zip(dets, g["dets"]):
            assert a[0] == b[0] and a[2] == b[2] and abs(a[1] - b[1]) < 1e-3 and abs(a[3] - b[3]) < 0.2
        ndet += len(dets)
    assert ndet >= 3


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(mg.DFT_WB_CASES))
def test_dft_wideband_cuda_reproduces_golden(cuda, name):
    from radiosonde_auto_rx_b200 import dft
    iq = mg.dft_wb_input(name)
    sr = mg.DFT_WB_CASES[name][0]
    fs = list(GW[name]["fq"])
    got = dft.DftWidebandBatch(sr, 15.0, True, nchan=len(fs)).run(iq, np.array([float(f) / sr for f in fs]), cont=True)
    for (rc, dets), f in zip(got, fs):
        g = GW[name]["fq"][f]
        assert rc == g["exit_code"] and len(dets) == len(g["dets"]), (f, rc, dets, g)
        for d, e in zip(dets, g["dets"]):
            assert d.type == e[0] and d.pos == e[2] and abs(d.score - e[1]) < 2e-4 and abs(d.df_hz - e[3]) < 0.2


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(mg.FSK_CASES))
def test_fsk_cuda_reproduces_golden(cuda, name):
    """The CUDA path straight against the reference's own outputs (no oracle in between)."""
    from radiosonde_auto_rx_b200 import fsk
    cfg = po.FskConfig(**mg.FSK_CASES[name][0])
    iq = mg.fsk_input(name)
    b = fsk.FskBatch(cfg.Fs, cfg.Rs, cfg.M, cfg.P, cfg.Nsym, 1000, cfg.fs_tx, cfg.lo, cfg.hi, 1 if cfg.mask else 0, nchan=1)
    out, frames = b.push(iq[None], "cs16", soft=cfg.soft, softinv=cfg.softinv)
    assert out[0, : frames[0] * b.Nbits].tobytes() == G[name + "/out"].tobytes()


@pytest.mark.gpu
def test_dft_cuda_reproduces_golden(cuda):
    from radiosonde_auto_rx_b200 import dft
    names = list(mg.DFT_CASES)
    iq = np.stack([mg.dft_input(n) for n in names])
    got = dft.DftDetectBatch(48000, 15.0, True, nchan=len(names)).run(iq, cont=True)
    for (rc, dets), n in zip(got, names):
        assert rc == GD[n]["exit_code"] and len(dets) == len(GD[n]["dets"])
        for d, e in zip(dets, GD[n]["dets"]):
            assert d.type == e[0] and d.pos == e[2] and abs(d.score - e[1]) < 2e-4
