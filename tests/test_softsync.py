"""soft-bit header sync + RS41 frame assembly.
CPU: the oracle restatement vs `oracle/_ref/rs41mod -r --softin -i` (raw frame hex).  GPU: IQ -> fsk_demod kernel -> soft bits
stay in HBM -> frame-sync kernel, compared frame for frame with the oracle and with the reference chain's stdout."""
import os

import numpy as np
import pytest

from oracle import pyoracle as po
from radiosonde_auto_rx_b200 import siggen

needs_ref = pytest.mark.skipif(not os.path.exists(os.path.join(po.REF_DIR, "rs41mod")), reason="oracle/_ref/rs41mod absent")


def soft_stream(seed, nsamp=400_000, ebno=12.0):
    iq = siggen.fsk_iq(seed, nsamp, ebno_db=ebno, ppm=12.0)
    out, _, _ = po.Oracle(po.RS41).run_stream(iq)
    return iq, out


@needs_ref
@pytest.mark.parametrize("seed,ebno", [(3, 14.0), (4, 9.0), (5, 6.0), (6, 30.0)])
def test_softsync_oracle_matches_reference_rs41mod(built, seed, ebno):
    _, soft = soft_stream(seed, ebno=ebno)
    ref = po.ref_cli_rs41mod_raw(soft.tobytes(), inv=True)
    mine = po.softsync_oracle(soft, False, True)
    assert len(ref) == len(mine) > 3
    for m, r in zip(mine, ref):
        assert "[NO]" in r[1]                       # random payload: the Reed-Solomon stage leaves the bytes alone
        n = min(m[4], m[0])                         # a frame cut short by EOF is only defined up to the bytes actually read
        assert m[3].tobytes().hex()[: 2 * n] == r[0][: 2 * n]


def test_softsync_polarity_and_softinv(built):
    _, soft = soft_stream(9)
    a = po.softsync_oracle(soft, False, True)
    b = po.softsync_oracle(-soft, True, True)        # --softinv undoes a sign flip
    c = po.softsync_oracle(soft, False, False)       # wrong polarity: the header correlates negatively -> nothing (aut == 0)
    assert len(a) == len(b) > 0 and all(x[3].tobytes() == y[3].tobytes() for x, y in zip(a, b))
    assert len(c) == 0


@pytest.mark.gpu
def test_fsk_demod_to_frames_on_device(cuda):
    import torch
    from radiosonde_auto_rx_b200 import fsk
    cfg = po.RS41
    nchan, nsamp = 6, 300_000
    iq = siggen.fsk_batch(3300, nchan, nsamp)
    iq[4] = siggen.noise_iq(8, nsamp)
    x = torch.from_numpy(iq.astype(np.float32) / np.float32(1000)).cuda()
    b = fsk.FskBatch(cfg.Fs, cfg.Rs, cfg.M, cfg.P, cfg.Nsym, 1000, cfg.fs_tx, cfg.lo, cfg.hi, 1, nchan=nchan)
    maxf = b.max_frames(nsamp)
    soft = torch.zeros((nchan, maxf * b.Nbits), dtype=torch.float32, device="cuda")
    b.demod_device(x, "f32", nsamp, soft, soft=True, stream=torch.cuda.current_stream())
    torch.cuda.synchronize()
    frames, _ = b.counts()
    nbits = torch.from_numpy((frames * b.Nbits).astype(np.int32)).cuda()
    got = fsk.rs41_framesync_device(soft, maxf * b.Nbits, nbits, nchan, softinv=False, inv=True)
    soft_h = soft.cpu().numpy()
    k = 0
    for c in range(nchan):
        exp = po.softsync_oracle(soft_h[c, : frames[c] * b.Nbits], False, True)
        mine = [g for g in got if g[0] == c]
        assert len(mine) == len(exp)
        for g, e in zip(mine, exp):
            assert (g[1], g[2], g[5]) == (e[0], e[1], e[4]), (c, g[:4], e[:3], g[5], e[4])
            assert g[4].tobytes() == e[3].tobytes(), (c, g[2], int(np.argmax(g[4] != e[3])))
            assert abs(g[3] - e[2]) < 1e-5
        k += len(mine)
        if c == 0 and os.path.exists(os.path.join(po.REF_DIR, "rs41mod")):          # and against the reference chain itself
            ref = po.ref_cli_rs41mod_raw(soft_h[c, : frames[c] * b.Nbits].tobytes(), inv=True)
            for r, g in zip(ref, mine):
                if "[NO]" not in r[1]:
                    continue        # the Reed-Solomon stage (not part of this row) rewrote the frame, e.g. a nearly empty EOF frame -> all zeros
                n = min(g[5], g[1])
                assert r[0][: 2 * n] == g[4].tobytes().hex()[: 2 * n]
            assert len(ref) == len(mine)
    assert k > 20
