"""GPU parity tests proper: the CUDA path through the C-ABI vs the oracle, bit for bit."""
import os
import numpy as np
import pytest

from oracle import pyoracle as po
from radiosonde_auto_rx_b200 import fsk, siggen

pytestmark = pytest.mark.gpu

CONFIGS = {
    # name: (FskConfig, siggen kwargs)   -- the auto_rx production lines (decode.py:869-1467)
    "rs41": (po.RS41, dict()),
    "default_hard": (po.FskConfig(48000, 4800, 2, 10, 50, soft=False), dict()),
    "rs92": (po.FskConfig(48000, 4800, 2, 10, 50, -20000, 20000), dict()),
    "dfm_softinv": (po.FskConfig(50000, 2500, 2, 10, 50, -5000, 5000, softinv=True), dict(Fs=50000, Rs=2500, spacing=5000.0)),
    "m10": (po.FskConfig(48080, 9616, 2, 5, 50, -10000, 10000), dict(Fs=48080, Rs=9616)),
    "meisei": (po.FskConfig(48000, 2400, 2, 10, 50, -15000, 15000), dict(Rs=2400)),
    "wxr301_mask": (po.FskConfig(96000, 4800, 2, 10, 50, -40000, 40000, 50000), dict(Fs=96000, Rs=4800, spacing=50000.0)),
    "4fsk_mask": (po.FskConfig(48000, 4800, 4, 10, 50, -20000, 20000, mask=4800), dict(M=4, spacing=4800.0)),
    "4fsk_peak_hard": (po.FskConfig(48000, 4800, 4, 10, 50, soft=False), dict(M=4, spacing=4800.0)),
}


def make_batch(cfg: po.FskConfig, nchan: int, device: int = 0, mode: str | None = None) -> fsk.FskBatch:
    """mode 'serial' forces the one-stage-at-a-time kernel; default lets the library pick (pipelined when eligible)."""
    import os
    old = os.environ.pop("B200FSK_KERNEL", None)
    if mode:
        os.environ["B200FSK_KERNEL"] = mode
    try:
        return fsk.FskBatch(cfg.Fs, cfg.Rs, cfg.M, cfg.P, cfg.Nsym, 1000, cfg.fs_tx, cfg.lo, cfg.hi, 1 if cfg.mask else 0,
                            nchan=nchan, device=device)
    finally:
        os.environ.pop("B200FSK_KERNEL", None)
        if old is not None:
            os.environ["B200FSK_KERNEL"] = old


def channels(seed, nchan, nsamp, kw):
    iq = siggen.fsk_batch(seed, nchan, nsamp, **kw)
    if nchan >= 3:
        iq[2] = siggen.noise_iq(seed + 99, nsamp)          # a noise-only channel: nin changes on most frames
    return iq


@pytest.mark.parametrize("mode", ["auto", "serial", "pipe", "wpc", "duo"])
@pytest.mark.parametrize("name", list(CONFIGS))
def test_push_matches_oracle(cuda, name, mode):
    cfg, kw = CONFIGS[name]
    nchan, nsamp = 5, 150_000
    iq = channels(100 + len(name), nchan, nsamp, kw)
    b = make_batch(cfg, nchan, mode=None if mode == "auto" else mode)
    b.enable_stats(True)
    out, frames = b.push(iq, "cs16", soft=cfg.soft, softinv=cfg.softinv)
    for c in range(nchan):
        orc = po.Oracle(cfg)
        ref, cons, recs = orc.run_stream(iq[c])
        assert frames[c] == len(recs) > 0
        assert out[c, : frames[c] * b.Nbits].tobytes() == ref.tobytes(), f"{name} ch{c}"
        st = b.stats(c)
        last = recs[-1]
        assert st.nin == last["nin_next"]
        assert np.float32(st.ppm).tobytes() == last["ppm"].tobytes()
        assert np.float32(st.norm_rx_timing).tobytes() == last["norm_rx_timing"].tobytes()
        assert np.float32(st.snr_est).tobytes() == last["snr_est"].tobytes()
        assert np.float32(st.EbNodB).tobytes() == last["EbNodB"].tobytes()
        assert np.array_equal(np.array(st.f_est, np.float32), last["f_est"][: cfg.M])
        assert np.float32(st.foff).tobytes() == last["foff"].tobytes()
        assert st.samples == cons
        assert b.Sf(c).tobytes() == orc.Sf().tobytes()


@pytest.mark.parametrize("mode", ["auto", "serial", "pipe", "wpc", "duo"])
def test_chunked_push_equals_single_push(cuda, mode):
    """Streaming: arbitrary chunking of the input must not change a single output bit (carry logic)."""
    cfg = po.RS41
    nchan, nsamp = 3, 100_000
    iq = channels(7, nchan, nsamp, {})
    ref = [po.Oracle(cfg).run_stream(iq[c])[0] for c in range(nchan)]
    b = make_batch(cfg, nchan, mode=None if mode == "auto" else mode)
    rng = np.random.default_rng(5)
    pos, got = 0, [[] for _ in range(nchan)]
    while pos < nsamp:
        n = int(min(nsamp - pos, rng.integers(1, 9000)))
        out, frames = b.push(iq[:, pos:pos + n], "cs16", soft=True)
        for c in range(nchan):
            got[c].append(out[c, : frames[c] * b.Nbits].copy())
        pos += n
    for c in range(nchan):
        assert np.concatenate(got[c]).tobytes() == ref[c].tobytes()


@pytest.mark.parametrize("mode", ["auto", "wpc", "duo"])
def test_max_frames_cap_and_stats_cadence(cuda, mode):
    cfg = po.RS41
    iq = channels(11, 1, 60_000, {})
    ref, _, recs = po.Oracle(cfg).run_stream(iq[0])
    b = make_batch(cfg, 1, mode=None if mode == "auto" else mode)
    b.enable_stats(True)
    got, k, pos = [], 0, 0
    while True:
        n = min(7000, 60_000 - pos)
        out, frames = b.push(iq[:, pos:pos + n], "cs16", soft=True, max_frames=2)
        pos += n
        got.append(out[0, : frames[0] * b.Nbits].copy())
        k += int(frames[0])
        if frames[0]:
            st = b.stats(0)
            assert np.float32(st.snr_est).tobytes() == recs[k - 1]["snr_est"].tobytes()
        if n == 0 and frames[0] == 0:
            break
    assert np.concatenate(got).tobytes() == ref.tobytes()


@pytest.mark.parametrize("mode", ["auto", "wpc", "duo"])
@pytest.mark.parametrize("fmt", ["f32", "cu8", "s16"])
def test_input_formats(cuda, fmt, mode):
    cfg = po.FskConfig(48000, 4800, 2, 10, 50, lower=(0 if fmt == "s16" else None))
    iq = siggen.fsk_iq(21, 80_000, ebno_db=12)
    if fmt == "f32":
        data = (iq.astype(np.float32) / np.float32(1000)).reshape(1, -1)
    elif fmt == "cu8":
        data = np.clip(iq.astype(np.int32) // 256 + 127, 0, 255).astype(np.uint8).reshape(1, -1)
    else:
        data = np.ascontiguousarray(iq[:, 0]).reshape(1, -1)
    ref, _, _ = po.Oracle(cfg).run_stream(data[0], fmt)
    b = make_batch(cfg, 1, mode=None if mode == "auto" else mode)
    out, frames = b.push(data, fmt, soft=True)
    assert frames[0] > 0 and out[0, : frames[0] * b.Nbits].tobytes() == ref.tobytes()


def test_device_resident_path(cuda):
    import torch
    cfg = po.RS41
    nchan, nsamp = 8, 120_000
    iq = channels(31, nchan, nsamp, {})
    x = torch.from_numpy(iq.astype(np.float32) / np.float32(1000)).cuda()       # exactly (float)int16/1000
    b = make_batch(cfg, nchan)
    maxf = b.max_frames(nsamp)
    out = torch.zeros((nchan, maxf * b.Nbits), dtype=torch.float32, device="cuda")
    b.demod_device(x, "f32", nsamp, out, soft=True, stream=torch.cuda.current_stream())
    torch.cuda.synchronize()
    frames, consumed = b.counts()
    o = out.cpu().numpy()
    for c in range(nchan):
        ref, cons, recs = po.Oracle(cfg).run_stream(iq[c])
        assert frames[c] == len(recs) and consumed[c] == cons
        assert o[c, : frames[c] * b.Nbits].tobytes() == ref.tobytes()
    assert b.last_kernel_ms() > 0


def test_reference_named_api(cuda):
    """fsk.h-style single-channel calls (fsk_create_hbr / fsk_nin / fsk_demod_sd)."""
    cfg = po.RS41
    iq = siggen.fsk_iq(41, 40_000, ebno_db=14).astype(np.float32) / np.float32(1000)
    orc = po.Oracle(cfg)
    m = fsk.fsk_create_hbr(cfg.Fs, cfg.Rs, cfg.M, cfg.P, cfg.Nsym, 1000, cfg.fs_tx)
    fsk.fsk_set_freq_est_limits(m, cfg.lo, cfg.hi)
    fsk.fsk_set_freq_est_alg(m, 1)
    pos = 0
    while pos + fsk.fsk_nin(m) <= len(iq):
        nin = fsk.fsk_nin(m)
        assert nin == orc.nin
        sd = fsk.fsk_demod_sd(m, iq[pos:pos + nin])
        assert sd.tobytes() == orc.frame(iq[pos:pos + nin]).tobytes()
        pos += nin
    st = fsk.fsk_get_demod_stats(m)
    assert np.float32(st.snr_est).tobytes() == np.float32(orc.last().snr_est).tobytes()
    fsk.fsk_destroy(m)


@pytest.mark.parametrize("mode", ["auto", "wpc", "duo"])
def test_nan_input_follows_reference_early_return(cuda, mode):
    """fsk.c:702-704: a NaN timing estimate returns early and the CLI re-emits its stale buffer."""
    cfg = po.FskConfig(48000, 4800, 2, 10, 50, softinv=True)
    x = (siggen.fsk_iq(51, 20_000, ebno_db=15).astype(np.float32) / np.float32(1000)).reshape(-1)
    x[2 * 2100] = np.nan
    ref, _, _ = po.Oracle(cfg).run_stream(x, "f32")
    b = make_batch(cfg, 1, mode=None if mode == "auto" else mode)
    out, frames = b.push(x.reshape(1, -1), "f32", soft=True, softinv=True)
    got = out[0, : frames[0] * b.Nbits]
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))


def test_exact_math_device_equals_libm(cuda):
    """norm_rx_timing of every frame (atan2f on the device) equals the oracle's libm atan2f bit for bit."""
    cfg = po.RS41
    iq = siggen.noise_iq(61, 200_000)
    ref, _, recs = po.Oracle(cfg).run_stream(iq)
    b = make_batch(cfg, 1)
    b.enable_stats(True)
    pos, k = 0, 0
    while pos < len(iq):
        out, frames = b.push(iq[None, pos:pos + 3100], "cs16", max_frames=1)
        pos += 3100
        if frames[0]:
            assert np.float32(b.stats(0).norm_rx_timing).tobytes() == recs[k]["norm_rx_timing"].tobytes()
            k += 1
    assert k > 50


def test_mixed_2fsk_4fsk_batch(cuda):
    """BASELINE.json configs[3] in miniature: channels at assorted rates / orders share the GPU, each bit-exact."""
    names = ["rs41", "rs92", "dfm_softinv", "m10", "4fsk_mask", "rs41", "meisei", "4fsk_mask", "wxr301_mask", "rs41"]
    cfgs, data, ref = [], [], []
    for i, n in enumerate(names):
        cfg, kw = CONFIGS[n]
        iq = siggen.fsk_iq(800 + i, 70_000, ebno_db=11.0, **kw)
        cfgs.append(dict(Fs=cfg.Fs, Rs=cfg.Rs, M=cfg.M, P=cfg.P, Nsym=cfg.Nsym, f1_tx=1000, fs_tx=cfg.fs_tx, est_min=cfg.lo,
                         est_max=cfg.hi, est_type=1 if cfg.mask else 0))
        data.append(iq)
        soft_cfg = po.FskConfig(cfg.Fs, cfg.Rs, cfg.M, cfg.P, cfg.Nsym, cfg.lower, cfg.upper, cfg.mask, True, False)
        ref.append(po.Oracle(soft_cfg).run_stream(iq)[0])
    mb = fsk.FskMixedBatch(cfgs)
    assert len(mb.batches) == 7
    got = mb.push(data, "cs16", soft=True)
    for c in range(len(names)):
        assert got[c].tobytes() == ref[c].tobytes(), names[c]
    mb.close()


@pytest.mark.parametrize("mode", ["auto", "wpc", "duo"])
def test_empty_and_sub_frame_pushes(cuda, mode):
    """Edge cases of the stream interface: zero-length pushes, pushes shorter than a frame (nothing may come out until a
    whole frame is available), and sample-by-sample feeding around a frame boundary."""
    cfg = po.RS41
    iq = siggen.fsk_iq(91, 12_500, ebno_db=12.0)
    ref, _, recs = po.Oracle(cfg).run_stream(iq)
    b = make_batch(cfg, 1, mode=None if mode == "auto" else mode)
    out, fr = b.push(iq[None, :0], "cs16")
    assert fr[0] == 0
    got = []
    cuts = [0, 1, 2, 2999, 3000, 3001, 3002, 5990, 6010, 9000, 9001, 12_500]
    for a, c in zip(cuts[:-1], cuts[1:]):
        out, fr = b.push(iq[None, a:c], "cs16")
        got.append(out[0, : fr[0] * b.Nbits].copy())
        out, fr = b.push(iq[None, :0], "cs16")            # an empty push in between changes nothing
        assert fr[0] == 0
    assert np.concatenate(got).tobytes() == ref.tobytes() and len(recs) == 4


@pytest.mark.parametrize("mode", ["pipe", "auto", "duo"])
def test_more_channels_than_resident_ctas(cuda, mode):
    """600 channels > 2 CTAs x 148 SMs: several waves of channel CTAs, every channel still bit-exact; ragged content
    (noise-only, different SNR / ppm / offsets) exercises different nin sequences per channel.  'auto' picks the
    many-channels-per-CTA throughput kernel for this batch size (16 channels share one lane-per-chain oscillator warp),
    'pipe' keeps the one-CTA-per-channel pipeline running in waves."""
    cfg = po.RS41
    nchan, nsamp = 600, 40_000
    iq = siggen.fsk_batch(5150, nchan, nsamp)
    iq[7] = siggen.noise_iq(1, nsamp); iq[311] = siggen.noise_iq(2, nsamp); iq[599] = siggen.noise_iq(3, nsamp)
    b = make_batch(cfg, nchan, mode=None if mode == "auto" else mode)
    out, frames = b.push(iq, "cs16")
    for c in list(range(0, nchan, 37)) + [7, 311, 599]:
        ref, _, recs = po.Oracle(cfg).run_stream(iq[c])
        assert frames[c] == len(recs) and out[c, : frames[c] * b.Nbits].tobytes() == ref.tobytes(), c


def test_full_ctas_of_the_throughput_kernel_two_waves(cuda):
    """2300 channels = 15 channel warps per CTA, more CTAs than SMs (two waves), the last CTA only partly filled; ragged content and
    per-channel stream lengths (device path with in_off): every checked channel bit-exact, and the kernel really is fsk_wpc_kernel."""
    import torch
    cfg = po.RS41
    nchan, nsamp = 2300, 30_000
    base = siggen.fsk_batch(4242, 64, nsamp)
    base[3] = siggen.noise_iq(9, nsamp)
    iq = np.tile(base, ((nchan + 63) // 64, 1, 1))[:nchan].copy()
    rng = np.random.default_rng(1)
    off = rng.integers(0, 2500, nchan).astype(np.int32)            # every channel starts somewhere else in its row
    off[:4] = 0
    b = make_batch(cfg, nchan)
    x = torch.from_numpy(iq).cuda()
    d_off = torch.from_numpy(off).cuda()
    maxf = b.max_frames(nsamp)
    d_out = torch.zeros((nchan, maxf * b.Nbits), dtype=torch.float32, device="cuda")
    b.demod_device(x, "cs16", nsamp, d_out, soft=True, in_stride=nsamp, out_stride=maxf * b.Nbits, d_in_off=d_off)
    frames, consumed = b.counts()
    assert b.last_kernel() == "fsk_wpc_kernel"
    out = d_out.cpu().numpy()
    for c in list(range(0, nchan, 61)) + [3, 67, 2285, 2299]:
        ref, cons, recs = po.Oracle(cfg).run_stream(iq[c, off[c]:])
        assert frames[c] == len(recs) and consumed[c] == cons, c
        assert out[c, : frames[c] * b.Nbits].tobytes() == ref.tobytes(), c


def test_nin_dependent_window_count_uses_the_serial_kernel(cuda):
    """Nsym=64 at Ts=10: floor(nin/64) is 9 for nin=635 and 10 for nin=645, so the number of FFT windows depends on nin and
    the pipelined kernel is not eligible -> the library must fall back to the serial kernel transparently."""
    cfg = po.FskConfig(48000, 4800, 2, 10, 64, -5000, 5000)
    iq = siggen.fsk_iq(77, 120_000, ebno_db=12.0)
    ref, _, recs = po.Oracle(cfg).run_stream(iq)
    b = make_batch(cfg, 2)
    out, fr = b.push(np.stack([iq, iq]), "cs16")
    assert fr[0] == fr[1] == len(recs) > 0
    assert out[0, : fr[0] * b.Nbits].tobytes() == ref.tobytes() and out[1, : fr[1] * b.Nbits].tobytes() == ref.tobytes()


def test_configuration_too_large_for_shared_memory_is_an_error(cuda):
    with pytest.raises(fsk.B200Error, match="error -3"):
        fsk.FskBatch(96000, 2400, 4, 10, 600, nchan=1)


@pytest.mark.parametrize("mode", ["wpc", "duo"])
def test_channels_that_end_at_different_samples_share_an_oscillator_warp(cuda, mode):
    """Regression: the oscillator warp steps all its chains in lockstep; the chain of a channel whose last frame has ended must stop
    writing that channel's ring (it is no longer covered by the ring-space check while the channel's consumers may still be in the
    last blocks of the last frame).  The race was timing dependent, so the push is repeated."""
    cfg, kw = CONFIGS["rs41"]
    nchan, nsamp = 5, 150_000
    iq = channels(104, nchan, nsamp, kw)
    ref_b = make_batch(cfg, nchan, mode="serial")
    ref, ref_frames = ref_b.push(iq, "cs16", soft=True)
    for rep in range(12):
        b = make_batch(cfg, nchan, mode=mode)
        out, frames = b.push(iq, "cs16", soft=True)
        assert np.array_equal(frames, ref_frames)
        for c in range(nchan):
            n = frames[c] * b.Nbits
            assert out[c, :n].tobytes() == ref[c, :n].tobytes(), f"repetition {rep} ch{c}"


def test_between_one_and_four_channels_per_sm_the_library_picks_the_duo_kernel(cuda):
    """300 channels on 148 SMs = two waves of two-channel CTAs, the last CTA of the second wave half filled; ragged starts."""
    import torch
    cfg = po.RS41
    nchan, nsamp = 2 * torch.cuda.get_device_properties(0).multi_processor_count + 5, 40_000
    base = siggen.fsk_batch(777, 32, nsamp)
    base[5] = siggen.noise_iq(11, nsamp)
    iq = np.tile(base, ((nchan + 31) // 32, 1, 1))[:nchan].copy()
    rng = np.random.default_rng(3)
    off = rng.integers(0, 2500, nchan).astype(np.int32)
    off[:2] = 0
    b = make_batch(cfg, nchan)
    x = torch.from_numpy(iq).cuda()
    d_off = torch.from_numpy(off).cuda()
    maxf = b.max_frames(nsamp)
    d_out = torch.zeros((nchan, maxf * b.Nbits), dtype=torch.float32, device="cuda")
    b.demod_device(x, "cs16", nsamp, d_out, soft=True, in_stride=nsamp, out_stride=maxf * b.Nbits, d_in_off=d_off)
    frames, consumed = b.counts()
    assert b.last_kernel() == "fsk_duo_kernel"
    out = d_out.cpu().numpy()
    for c in list(range(0, nchan, 23)) + [1, 5, 37, nchan - 2, nchan - 1]:
        ref, cons, recs = po.Oracle(cfg).run_stream(iq[c, off[c]:])
        assert frames[c] == len(recs) and consumed[c] == cons, c
        assert out[c, : frames[c] * b.Nbits].tobytes() == ref.tobytes(), c


@pytest.mark.parametrize("name", list(CONFIGS))
def test_every_production_line_against_the_unmodified_reference_library(cuda, name):
    """The same comparison as test_push_matches_oracle, but straight against oracle/_ref/libfsk_ref.so (utils/fsk.c + kiss_fft compiled
    unmodified with the reference's flags) instead of the C restatement: output bits, consumed samples and the carried spectrum Sf."""
    if not os.path.exists(os.path.join(po.REF_DIR, "libfsk_ref.so")):
        pytest.skip("oracle/_ref/libfsk_ref.so absent")
    cfg, kw = CONFIGS[name]
    nchan, nsamp = 3, 60_000
    iq = channels(300 + len(name), nchan, nsamp, kw)
    b = make_batch(cfg, nchan)
    b.enable_stats(True)
    out, frames = b.push(iq, "cs16", soft=cfg.soft, softinv=cfg.softinv)
    for c in range(nchan):
        ref, cons, recs, Sf = po.ref_run_stream(cfg, iq[c])
        assert frames[c] == len(recs) > 0
        assert out[c, : frames[c] * b.Nbits].tobytes() == ref.tobytes(), f"{name} ch{c}"
        assert b.stats(c).samples == cons
        mine = b.Sf(c)
        assert mine.tobytes() == Sf[: mine.size].tobytes()
