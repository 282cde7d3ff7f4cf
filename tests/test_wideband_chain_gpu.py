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


def test_streaming_iq_to_corrected_frames(cuda):
    """The auto_rx RS41 chain as a stream, device-resident: IQ arrives in blocks -> fsk_demod (state in the handle, cursor per channel)
    -> soft bits stay in HBM -> streaming frame sync + Reed-Solomon (--ecc2).  Well-formed frames are transmitted, so the decoded,
    corrected frames must be the transmitted ones; and the whole stream must equal the oracle chain run over the complete IQ."""
    import os
    import sys
    import torch
    from radiosonde_auto_rx_b200 import fsk
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import rs41_frames as rf
    cfg = po.RS41
    nchan, nfr = 3, 6
    rng = np.random.default_rng(77)
    clean, iq = [], []
    nsamp = (nfr * 4800 + 600) * 10
    for c in range(nchan):
        frames = [rf.clean_frame(rng, frnb=100 * c + k) for k in range(nfr)]
        bits = np.concatenate([rng.integers(0, 2, 300, dtype=np.uint8)] + [np.concatenate([rf.frame_bits(f), rng.integers(0, 2, 2240, dtype=np.uint8)]) for f in frames])
        clean.append(frames)
        iq.append(siggen.fsk_iq(900 + c, nsamp, ebno_db=(14.0, 11.0, 9.5)[c], ppm=(-6.0, 0.0, 9.0)[c], bits=bits, header=None))
    iq = np.stack(iq)
    x = torch.from_numpy(iq.astype(np.float32) / np.float32(1000)).cuda()
    b = fsk.FskBatch(cfg.Fs, cfg.Rs, cfg.M, cfg.P, cfg.Nsym, 1000, cfg.fs_tx, cfg.lo, cfg.hi, 1, nchan=nchan)
    sync = fsk.Rs41FrameSync(nchan, ecc=2)
    off = torch.zeros(nchan, dtype=torch.int32, device="cuda")
    step = 37000
    maxf = b.max_frames(nsamp)                       # the C-ABI sizes rows by the absolute sample count, whatever the cursors are
    soft = torch.zeros((nchan, maxf * b.Nbits), dtype=torch.float32, device="cuda")
    got, all_soft = [], [[] for _ in range(nchan)]
    for avail in list(range(step, nsamp, step)) + [nsamp]:
        b.demod_device(x, "f32", avail, soft, soft=True, d_in_off=off, stream=torch.cuda.current_stream())
        torch.cuda.synchronize()
        frames, consumed = b.counts()
        off += torch.from_numpy(consumed.astype(np.int32)).cuda()
        nbits = torch.from_numpy((frames * b.Nbits).astype(np.int32)).cuda()
        got += sync.push(soft, maxf * b.Nbits, nbits)
        sh = soft.cpu().numpy()
        for c in range(nchan):
            all_soft[c].append(sh[c, : frames[c] * b.Nbits].copy())
    got += sync.push(soft, maxf * b.Nbits, 0, flush=True)
    ok_total = 0
    for c in range(nchan):
        exp_soft, _, _ = po.Oracle(cfg).run_stream(iq[c].astype(np.float32) / np.float32(1000), "f32")
        mine_soft = np.concatenate(all_soft[c])
        assert mine_soft.tobytes() == exp_soft.tobytes(), f"channel {c}: blockwise demodulation differs from one pass of the oracle"
        exp = po.softsync_oracle(exp_soft, False, True, ecc=2)
        mine = sorted([g for g in got if g[0] == c], key=lambda g: g[2])
        assert len(mine) == len(exp) >= nfr - 1
        for g, e in zip(mine, exp):
            assert (g[1], g[2], g[5], g[6], g[7]) == (e[0], e[1], e[4], e[5], e[6]) and g[4].tobytes() == e[3].tobytes()
        good = [g for g in mine if g[6] >= 0 and g[7] == 0x3F]
        for g in good:
            assert any(np.array_equal(g[4][:320], f) for f in clean[c]), "a frame that passed RS + all CRCs is not one of the transmitted frames"
        ok_total += len(good)
    assert ok_total >= 2 * (nfr - 1)
