"""GPU parity: batched dft_detect (C-ABI) vs the oracle: identical type / sign / sample position / exit code,
score within 2e-4 abs (the reference itself is -Ofast: float sums are not bit-reproducible), df within 0.2 Hz."""
import numpy as np
import pytest

from oracle import pyoracle as po
from radiosonde_auto_rx_b200 import dft, siggen

pytestmark = pytest.mark.gpu
SCORE_TOL, DF_TOL = 2e-4, 0.2


def compare(got, exp_rc, exp_dets, tag=""):
    rc, dets = got
    assert rc == exp_rc, f"{tag}: exit code {rc} != {exp_rc}"
    assert len(dets) == len(exp_dets), f"{tag}: {len(dets)} detections, oracle {len(exp_dets)}"
    for d, e in zip(dets, exp_dets):
        assert d.type == e[0] and d.pos == e[2], f"{tag}: {d} vs {e}"
        assert abs(d.score - e[1]) < SCORE_TOL and (d.score < 0) == (e[1] < 0), f"{tag}: score {d.score} vs {e[1]}"
        assert abs(d.df_hz - e[3]) < DF_TOL


def build_channels(nsamp, kinds, seed=300):
    chans, tags = [], []
    for i, k in enumerate(kinds):
        if k == "noise":
            chans.append(siggen.noise_iq(seed + i, nsamp))
        else:
            tail = ""
            if k == "M10":          # frame-type bits after the header decide M10 vs M20 (dft_detect.c:1526-1540)
                tail = "0110010010011111" if i % 2 == 0 else "0100010100100000"
                tail = "".join("10" if b == "1" else "01" for b in tail)
            chans.append(siggen.fm_header_iq(seed + i, nsamp, k, snr_db=(12, 18, 25)[i % 3], invert=(i % 4 == 3), tail=tail))
        tags.append(k)
    return np.stack(chans), tags


@pytest.mark.parametrize("cont", [True, False])
def test_header_bank_matches_oracle(cuda, cont):
    kinds = list(siggen.HEADER_BANK) + ["noise", "M10", "RS41"]
    nsamp = 100_000
    iq, tags = build_channels(nsamp, kinds)
    b = dft.DftDetectBatch(48000, 15.0, True, nchan=len(kinds))
    got = b.run(iq, cont=cont)
    found = 0
    for c, k in enumerate(kinds):
        rc, dets, _ = po.dft_oracle_run(iq[c], 48000, 15.0, True, cont)
        compare(got[c], rc, dets, f"{k}[{c}] cont={cont}")
        found += len(dets) > 0
    assert found >= len(siggen.HEADER_BANK)
    fr, co = b.kernel_ms()
    assert fr > 0 and co > 0 and b.launches == 3


def test_time_limit_and_device_path(cuda):
    import torch
    kinds = ["RS41", "DFM9", "noise", "M10"]
    iq, _ = build_channels(160_000, kinds, seed=500)
    b = dft.DftDetectBatch(48000, 15.0, True, nchan=len(kinds))
    d = torch.from_numpy(iq).cuda()
    got = b.run_device(d, "cs16", iq.shape[1], cont=True, tl=1.5)
    for c, k in enumerate(kinds):
        rc, dets, _ = po.dft_oracle_run(iq[c], 48000, 15.0, True, True, 1.5)
        compare(got[c], rc, dets, f"{k} tl=1.5")


def test_no_dc_and_other_bandwidth(cuda):
    kinds = ["RS41", "MEISEI", "LMS6"]
    iq, _ = build_channels(90_000, kinds, seed=700)
    for bw, dc in ((20.0, True), (15.0, False)):
        b = dft.DftDetectBatch(48000, bw, dc, nchan=len(kinds))
        got = b.run(iq, cont=True)
        for c, k in enumerate(kinds):
            rc, dets, _ = po.dft_oracle_run(iq[c], 48000, bw, dc, True)
            compare(got[c], rc, dets, f"{k} bw={bw} dc={dc}")


@pytest.mark.parametrize("sr,bw,cont,tl", [(48000, 15.0, True, -1.0), (48000, 15.0, False, -1.0), (96000, 64.0, True, -1.0), (48000, 15.0, True, 2.0)])
def test_streaming_pushes_equal_whole_buffer(cuda, sr, bw, cont, tl):
    """b200dft_stream_*: arbitrary chunking of the input (1 sample .. several blocks per push) gives the detections, positions and exit
    codes of one whole-buffer run -- including channels whose iMet AFSK branch shifts the block schedule -- while the handle only keeps
    the tail later blocks can reach."""
    nsamp = 300_000 if sr == 48000 else 400_000
    kinds = ["RS41", "DFM9", "MRZ", "noise", "MEISEI", "M10"] if sr == 48000 else ["WXR301", "RS41", "noise", "DFM9"]
    chans = []
    for i, k in enumerate(kinds):
        if k == "noise":
            chans.append(siggen.noise_iq(40 + i, nsamp))
        else:
            chans.append(siggen.fm_header_iq(9000 + i, nsamp, k, Fs=sr, dev_hz=12000.0 if k.startswith("WXR") else 2400.0, snr_db=14.0, invert=(i == 2)))
    iq = np.stack(chans)
    b = dft.DftDetectBatch(sr, bw, True, nchan=len(kinds))
    whole = b.run(iq, cont=cont, tl=tl)
    b.stream_begin("cs16", cont=cont, tl=tl)
    rng = np.random.default_rng(7)
    pos, got = 0, None
    while pos < nsamp:
        n = int(min(nsamp - pos, rng.choice([1, 37, 1500, 4000, 11000, 30000])))
        got = b.stream_push(iq[:, pos:pos + n])
        pos += n
    got = b.stream_push(iq[:, :0], flush=True)
    for c, k in enumerate(kinds):
        rc, dets = whole[c]
        assert got[c][0] == rc, (k, got[c][0], rc)
        g = got[c][1][:len(dets)]          # the whole-buffer call keeps at most B200DFT_MAXDET detections, the stream reports all
        assert len(got[c][1]) >= len(dets) and [(d.type, d.pos) for d in g] == [(d.type, d.pos) for d in dets], k
        for a, e in zip(g, dets):
            assert abs(a.score - e.score) < 1e-6 and abs(a.df_hz - e.df_hz) < 1e-3
    assert sum(len(g[1]) for g in got) >= 3


def test_output_text_matches_reference_format(cuda):
    iq, _ = build_channels(60_000, ["RS41"], seed=900)
    b = dft.DftDetectBatch(48000, 15.0, True, nchan=1)
    rc, dets = b.run(iq, cont=False)[0]
    txt = b.format_lines(dets)
    typ, rest = txt.strip().split(":")
    assert typ == "RS41" and rest.strip().endswith("Hz") and rc == 3


def test_imet_afsk_branch_shifts_the_block_schedule(cuda):
    """MRZ / MEISEI style 2400-baud headers also trip the iMet AFSK preamble correlator; the reference then swallows one
    extra second of samples outside its block counter (dft_detect.c:1542-1607), which moves every later block.  Long streams
    must still agree detection for detection."""
    nsamp = 1 << 18
    chans = [siggen.fm_header_iq(9001, nsamp, "RS41", snr_db=14.0), siggen.fm_header_iq(9000, nsamp, "DFM9", snr_db=14.0),
             siggen.fm_header_iq(9009, nsamp, "MRZ", snr_db=10.0, invert=True), siggen.noise_iq(4, nsamp),
             siggen.fm_header_iq(9007, nsamp, "MEISEI", snr_db=14.0), siggen.fm_header_iq(9025, nsamp, "MRZ", snr_db=20.0)]
    iq = np.stack(chans)                    # shifted channels sit at indices 2, 4, 5: the relaunch must map them back correctly
    b = dft.DftDetectBatch(48000, 15.0, True, nchan=len(chans))
    got = b.run(iq, cont=True)
    shifted = 0
    for c in range(len(chans)):
        rc, dets, blks = po.dft_oracle_run(iq[c], 48000, 15.0, True, True, want_blocks=True)
        compare(got[c], rc, dets[:64], f"ch{c}")
        shifted += len(blks) < 14 * (nsamp // (b.K - 4))
    assert shifted >= 1, "the test signal no longer exercises the iMet branch"
    assert b.launches > 3          # at least one relaunch happened


# ---- wideband mode: dft_detect --IQ <fq> --------------------------------------------------------------
WB_SR = 2400000
_M10_TAIL = "".join("10" if b == "1" else "01" for b in "0110010010011111")     # Manchester frame-type bits, as in build_channels
WB_EMIT = [("RS41", 300e3), ("DFM9", -500e3), ("M10", 800e3, _M10_TAIL), ("RS92", -900e3), ("LMS6", 1000e3), ("IMET5", -50e3)]


def _wb_capture(seed=1, secs=1.5, sr=WB_SR, snr_db=5.0, dc=(300.0, -200.0)):
    return siggen.wideband_mix(seed, int(sr * secs), sr, WB_EMIT, snr_db=snr_db, dc=dc)


def test_channeliser_matches_oracle_if_stream():
    """rotate + decimating FIR output (float IQ at the IF rate) against the restated dft_detect.c:751-761 front."""
    import torch
    from radiosonde_auto_rx_b200.dft import DftWidebandBatch
    wide = _wb_capture(secs=0.5)
    fqs = np.array([300e3, -500e3, 800e3, -900e3, 123456.0, 0.0, -1199999.0, 1200000.0]) / WB_SR
    b = DftWidebandBatch(WB_SR, 15.0, True, nchan=len(fqs))
    assert (b.sample_rate, b.decM, b.dectaps) == (48000, 50, 343)
    d_w = torch.from_numpy(wide).cuda()
    nout = wide.shape[0] // b.decM
    d_o = torch.zeros((len(fqs), nout, 2), dtype=torch.float32, device="cuda")
    assert b.channelise(d_w, "cs16", wide.shape[0], fqs, d_o) == nout
    got = d_o.cpu().numpy().view(np.complex64)[..., 0]
    for c, fq in enumerate(fqs):
        _, _, info, ref = po.dft_oracle_run_wideband(wide, WB_SR, float(fq), 15.0, True, True, want_if=True)
        assert info["if_n"] == nout
        err = np.abs(got[c] - ref).max()
        assert err <= 2e-6 * max(1.0, np.abs(ref).max()), (c, fq, err)        # float32 FIR, different summation order


@pytest.mark.parametrize("fmt", ["cs16", "f32", "cu8"])
def test_wideband_detect_matches_oracle(fmt):
    """type / sign / position / exit code identical, score and df within the --iq mode's tolerances, per candidate fq"""
    from radiosonde_auto_rx_b200.dft import DftWidebandBatch
    wide = _wb_capture(seed=3)
    if fmt == "f32":
        wide = (wide.astype(np.float32) / np.float32(32768.0))
    elif fmt == "cu8":
        wide = np.clip(np.rint(wide.astype(np.float32) / 96.0) + 128, 0, 255).astype(np.uint8)
    hz = [e[1] for e in WB_EMIT] + [100e3, -700e3]                      # two empty frequencies
    fqs = np.array(hz) / WB_SR
    b = DftWidebandBatch(WB_SR, 15.0, True, nchan=len(fqs))
    got = b.run(wide, fqs, cont=True)
    nfound = 0
    for c, fq in enumerate(fqs):
        rc, dets, _, _ = po.dft_oracle_run_wideband(wide, WB_SR, float(fq), 15.0, True, True)
        compare(got[c], rc, dets, f"{fmt} fq={fq:+.5f}")
        nfound += len(dets) > 0
    assert nfound >= len(WB_EMIT) - 1


def test_wideband_odd_rate_and_time_limit():
    """sr_base that 48000 does not divide: the IF rate walks up to the next divisor (dft_detect.c:1038-1041); -t stops early"""
    from radiosonde_auto_rx_b200.dft import DftWidebandBatch
    sr = 1000000                                                         # -> IF 50000, decM 20
    wide = siggen.wideband_mix(5, int(sr * 2.5), sr, [("RS41", 200e3), ("DFM9", -300e3)], snr_db=8.0)
    fqs = np.array([200e3, -300e3, 0.0]) / sr
    b = DftWidebandBatch(sr, 15.0, True, nchan=3)
    assert (b.sample_rate, b.decM) == (50000, 20)
    for cont, tl in ((False, -1.0), (True, 1.0)):
        got = b.run(wide, fqs, cont=cont, tl=tl)
        for c, fq in enumerate(fqs):
            rc, dets, _, _ = po.dft_oracle_run_wideband(wide, sr, float(fq), 15.0, True, cont, tl)
            assert got[c][0] == rc and [(d.type, d.pos) for d in got[c][1]] == [(o[0], o[2]) for o in dets], (cont, tl, c, got[c], dets)
