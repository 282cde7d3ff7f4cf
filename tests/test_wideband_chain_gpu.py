"""BASELINE cfg5 in small: ONE wideband capture -> channeliser -> per-channel fsk_demod -> RS41 frame sync (8f-1),
every intermediate staying in HBM.  Each stage is checked against its oracle on the exact input the previous GPU stage produced
(so the bit-exact bars of the single stages carry over), and the channeliser output itself against the restated reference front."""
import numpy as np
import pytest

from oracle import pyoracle as po
from radiosonde_auto_rx_b200 import siggen

pytestmark = pytest.mark.gpu
SR = 2400000


def capture(nsamp, centres, seed=40):
    acc = np.zeros((nsamp, 2), np.int32)
    for i, fc in enumerate(centres):           # RS41-like 2FSK, 4800 baud, tones at fc -/+ 2400 Hz, header every 4800 bits
        acc += siggen.fsk_iq(seed + i, nsamp, Fs=SR, Rs=4800, f_lo=fc - 2400.0, spacing=4800.0, ebno_db=22.0, amp=1500.0, ppm=(-8.0, 0.0, 11.0)[i % 3])
    return np.clip(acc, -32768, 32767).astype(np.int16)


def test_wideband_to_rs41_frames_on_device(cuda):
    import torch
    from radiosonde_auto_rx_b200 import fsk
    from radiosonde_auto_rx_b200.dft import DftWidebandBatch
    centres = [310e3, -455e3, 900e3]
    cand = centres + [120e3]                                             # one empty candidate
    nwide = int(SR * 2.0)
    wide = capture(nwide, centres)
    ch = DftWidebandBatch(SR, 15.0, True, nchan=len(cand))
    assert ch.sample_rate == 48000
    nout = nwide // ch.decM
    d_w = torch.from_numpy(wide).cuda()
    d_if = torch.zeros((len(cand), nout, 2), dtype=torch.float32, device="cuda")
    fqs = np.array(cand) / SR
    assert ch.channelise(d_w, "cs16", nwide, fqs, d_if) == nout
    # stage 1 against the restated reference front
    if_h = d_if.cpu().numpy()
    _, _, _, ref_if = po.dft_oracle_run_wideband(wide, SR, float(fqs[0]), 15.0, True, True, want_if=True)
    assert np.abs(if_h[0].view(np.complex64)[:, 0] - ref_if).max() <= 2e-6 * np.abs(ref_if).max()
    # stage 2: fsk_demod on the device-resident IF streams (the RS41 line of auto_rx, decode.py:901)
    cfg = po.RS41
    b = fsk.FskBatch(cfg.Fs, cfg.Rs, cfg.M, cfg.P, cfg.Nsym, 1000, cfg.fs_tx, cfg.lo, cfg.hi, 1, nchan=len(cand))
    maxf = b.max_frames(nout)
    soft = torch.zeros((len(cand), maxf * b.Nbits), dtype=torch.float32, device="cuda")
    b.demod_device(d_if, "f32", nout, soft, soft=True, stream=torch.cuda.current_stream())
    torch.cuda.synchronize()
    frames, _ = b.counts()
    soft_h = soft.cpu().numpy()
    for c in range(len(cand)):
        exp, _, _ = po.Oracle(cfg).run_stream(np.ascontiguousarray(if_h[c]), "f32")
        assert frames[c] * b.Nbits == exp.size
        assert soft_h[c, : exp.size].tobytes() == exp.tobytes(), f"candidate {c}: soft bits differ from the oracle on the same IF stream"
    # stage 3: header sync + frame assembly, still on the device
    nbits = torch.from_numpy((frames * b.Nbits).astype(np.int32)).cuda()
    got = fsk.rs41_framesync_device(soft, maxf * b.Nbits, nbits, len(cand), softinv=False, inv=True)
    per = [sum(1 for g in got if g[0] == c) for c in range(len(cand))]
    for c in range(len(cand)):
        exp = po.softsync_oracle(soft_h[c, : frames[c] * b.Nbits], False, True)
        mine = [g for g in got if g[0] == c]
        assert len(mine) == len(exp)
        for g, e in zip(mine, exp):
            assert (g[1], g[2]) == (e[0], e[1]) and g[4].tobytes() == e[3].tobytes()
    assert all(p >= 1 for p in per[:3]) and per[3] == 0, per          # 2 s of 4800 baud, header every second: >= 1 frame per sonde, none in the gap
