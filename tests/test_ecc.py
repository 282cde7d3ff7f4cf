"""the Reed-Solomon stage of rs41mod (two interleaved RS(255,231) words per frame, retry with the fixed packet ids).
CPU: oracle/ecc_oracle.c against the reference module itself (oracle/_ref/libbch_ref.so, word by word, also beyond t errors) and
against `oracle/_ref/rs41mod -r [--ecc2] --softin -i` frame by frame.  GPU: the same stage inside the frame-sync kernel."""
import os
import sys

import numpy as np
import pytest

from oracle import pyoracle as po

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import rs41_frames as rf  # noqa: E402

needs_bch = pytest.mark.skipif(not os.path.exists(os.path.join(po.REF_DIR, "libbch_ref.so")), reason="oracle/_ref/libbch_ref.so absent")
needs_ref = pytest.mark.skipif(not os.path.exists(os.path.join(po.REF_DIR, "rs41mod")), reason="oracle/_ref/rs41mod absent")


def test_crc16_known_answer(built):
    # the reference hard-codes the CRC of the standard zero block 76 11 00..00 as EC C7 (rs41mod.c:1757-1758)
    assert po.crc16(bytes(17)) == 0xC7EC
    assert po.crc16(b"123456789") == 0x29B1                      # CRC-16/CCITT-FALSE check value


def test_rs255_roundtrip(built):
    rng = np.random.default_rng(1)
    for nerr in (0, 1, 5, 11, 12):
        msg = rng.integers(0, 256, 231, dtype=np.uint8)
        cw = po.rs255_encode(msg)
        assert po.rs255_decode(cw)[0] == 0
        bad = cw.copy()
        idx = rng.choice(255, nerr, replace=False)
        bad[idx] ^= rng.integers(1, 256, nerr, dtype=np.uint8)
        rc, fixed = po.rs255_decode(bad)
        assert rc == nerr and np.array_equal(fixed, cw)


@needs_bch
def test_rs255_matches_reference_module(built):
    """same return code and same output word as bch_ecc_mod.c's rs_decode for 0..30 symbol errors (miscorrections included)"""
    enc, dec = po.ref_rs255()
    rng = np.random.default_rng(2)
    seen = set()
    for trial in range(600):
        msg = rng.integers(0, 256, 231, dtype=np.uint8)
        cw = po.rs255_encode(msg)
        assert np.array_equal(cw, enc(np.concatenate([np.zeros(24, np.uint8), msg])))
        nerr = int(rng.integers(0, 31)) if trial % 3 else int(rng.integers(11, 16))
        bad = cw.copy()
        idx = rng.choice(255, nerr, replace=False)
        bad[idx] ^= rng.integers(1, 256, nerr, dtype=np.uint8)
        if trial % 50 == 7:
            bad[:] = rng.integers(0, 256, 255, dtype=np.uint8)    # no codeword structure at all
        r_rc, r_cw = dec(bad)
        o_rc, o_cw = po.rs255_decode(bad)
        assert o_rc == r_rc, (trial, nerr, o_rc, r_rc)
        assert np.array_equal(o_cw, r_cw), (trial, nerr)
        seen.add(r_rc if r_rc < 0 else 0)
    assert {-1, 0} <= seen                                       # the failure exits were exercised (-2/-3 are rare)


@needs_ref
@pytest.mark.parametrize("ecc", [1, 2])
@pytest.mark.parametrize("seed,sigma,burst", [(1, 0.30, None), (2, 0.40, None), (3, 0.45, None), (4, 0.50, None), (5, 0.38, (2, 0x39, 2)),
                                               (6, 0.36, (1, 0x65, 30)), (7, 0.2, (3, 200, 40))])
def test_rs41_ecc_oracle_matches_reference_rs41mod(built, ecc, seed, sigma, burst):
    soft, clean = rf.soft_stream(seed, 8, sigma, burst=burst)
    ref = po.ref_cli_rs41mod_raw(soft.tobytes(), inv=True, ecc=ecc)
    mine = po.softsync_oracle(soft, False, True, ecc=ecc)
    assert len(ref) == len(mine) >= 6
    nok = 0
    for m, (hexs, tag) in zip(mine, ref):
        ec = m[5]
        exp_tag = "[OK]" + (f" ({ec})" if ec > 0 else "") if ec >= 0 else "[NO] " + {-1: "(-+)", -2: "(+-)", -3: "(--)"}[ec]
        assert tag.strip() == exp_tag, (tag, exp_tag)
        assert m[3].tobytes().hex() == hexs
        nok += ec >= 0
    if sigma <= 0.30 and burst is None:
        assert nok == len(mine) == len(clean)                    # and the corrected frames are the transmitted ones
        for m, c in zip(mine, clean):
            assert np.array_equal(m[3][:320], c)


def test_crc_mask_and_clean_frames(built):
    soft, clean = rf.soft_stream(11, 5, 0.25)
    fr = po.softsync_oracle(soft, False, True, ecc=2)
    assert len(fr) == 5
    for m, c in zip(fr, clean):
        assert m[5] >= 0 and m[6] == 0x3F and np.array_equal(m[3][:320], c)
    soft, _ = rf.soft_stream(12, 5, 0.55)                          # hopeless: nothing decodes, and no block CRC survives
    fr = po.softsync_oracle(soft, False, True, ecc=2)
    assert all(m[5] < 0 for m in fr) and all(m[6] != 0x3F for m in fr)


import hashlib  # noqa: E402
import importlib.util  # noqa: E402
import json  # noqa: E402

_spec = importlib.util.spec_from_file_location("make_golden", os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "make_golden.py"))
_mg = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_mg)
CASES = _mg.RS41_CASES
GR = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_rs41.json")))


def _tag(ec):
    return ("[OK]" + (f" ({ec})" if ec > 0 else "")) if ec >= 0 else "[NO] " + {-1: "(-+)", -2: "(+-)", -3: "(--)"}[ec]


@pytest.mark.parametrize("ecc", [1, 2])
@pytest.mark.parametrize("case", CASES, ids=lambda c: f"seed{c[0]}")
def test_rs41_oracle_reproduces_golden(built, case, ecc):
    """recorded output of the reference binary (tests/golden/golden_rs41.json): works without oracle/_ref"""
    seed, sigma, burst = case
    soft, _ = rf.soft_stream(seed, 8, sigma, burst=burst)
    g = GR[str(seed)]
    assert hashlib.sha256(soft.tobytes()).hexdigest() == g["sha"]
    mine = po.softsync_oracle(soft, False, True, ecc=ecc)
    assert [[hashlib.sha1(m[3].tobytes().hex().encode()).hexdigest()[:16], _tag(m[5])] for m in mine] == g[f"ecc{ecc}"]


@pytest.mark.gpu
@pytest.mark.parametrize("ecc", [1, 2])
def test_rs41_cuda_reproduces_golden(cuda, ecc):
    """the kernel straight against the reference's recorded frames (no oracle in between)"""
    import torch
    from radiosonde_auto_rx_b200 import fsk
    streams = [rf.soft_stream(s, 8, sg, burst=b)[0] for s, sg, b in CASES]
    n = max(s.size for s in streams)
    buf = np.zeros((len(streams), n), np.float32)
    for i, s in enumerate(streams):
        buf[i, : s.size] = s
    nb = torch.tensor([s.size for s in streams], dtype=torch.int32, device="cuda")
    got = fsk.rs41_framesync_device(torch.from_numpy(buf).cuda(), n, nb, len(streams), ecc=ecc)
    for c, (seed, _, _) in enumerate(CASES):
        mine = [[hashlib.sha1(g[4].tobytes().hex().encode()).hexdigest()[:16], _tag(g[6])] for g in got if g[0] == c]
        assert mine == GR[str(seed)][f"ecc{ecc}"], (seed, ecc)


@pytest.mark.gpu
@pytest.mark.parametrize("ecc", [1, 2])
def test_rs41_ecc_on_device(cuda, ecc):
    """frame sync + Reed-Solomon + CRC in one kernel pass over soft bits in HBM: every frame identical (bytes, return code, CRC mask)
    to the oracle, and to what the reference binary prints"""
    import torch
    from radiosonde_auto_rx_b200 import fsk
    streams = [rf.soft_stream(s, 8, sg, burst=b)[0] for s, sg, b in CASES]
    streams[3] = streams[3][: streams[3].size - 3000]              # one stream ends inside a frame (EOF rule, rs41mod.c:2481-2484)
    n = max(s.size for s in streams)
    buf = np.zeros((len(streams), n), np.float32)
    for i, s in enumerate(streams):
        buf[i, : s.size] = s
    d = torch.from_numpy(buf).cuda()
    nb = torch.tensor([s.size for s in streams], dtype=torch.int32, device="cuda")
    got = fsk.rs41_framesync_device(d, n, nb, len(streams), softinv=False, inv=True, ecc=ecc)
    total = rescued = 0
    for c, s in enumerate(streams):
        exp = po.softsync_oracle(s, False, True, ecc=ecc)
        mine = [g for g in got if g[0] == c]
        assert len(mine) == len(exp)
        for g, e in zip(mine, exp):
            assert (g[1], g[2], g[5], g[6], g[7]) == (e[0], e[1], e[4], e[5], e[6]), (c, g[:3], g[5:], e[:2], e[4:])
            assert g[4].tobytes() == e[3].tobytes(), (c, g[2])
        if os.path.exists(os.path.join(po.REF_DIR, "rs41mod")):
            ref = po.ref_cli_rs41mod_raw(s.tobytes(), inv=True, ecc=ecc)
            assert len(ref) == len(mine)
            for (hexs, tag), g in zip(ref, mine):
                assert g[4].tobytes().hex() == hexs and tag.strip().startswith("[OK]" if g[6] >= 0 else "[NO]")
        total += len(mine)
        rescued += sum(1 for g in mine if g[6] > 0)
    assert total >= 70 and rescued >= 20


@pytest.mark.gpu
@pytest.mark.parametrize("ecc", [0, 2])
def test_rs41_streaming_pushes_equal_whole_buffer(cuda, ecc):
    """the same streams cut into pushes of arbitrary sizes (1 bit ... several frames, different per channel) through the streaming
    handle give exactly the frames of the whole-buffer call (and therefore of the oracle / reference): state carried across pushes is
    the header ring, the frame in progress with its unfinished byte, and the frame buffer the EOF rule falls back on"""
    import torch
    from radiosonde_auto_rx_b200 import fsk
    streams = [rf.soft_stream(s, 8, sg, burst=b)[0] for s, sg, b in CASES[:7]]
    streams[3] = streams[3][: streams[3].size - 3000]
    streams[5] = streams[5][: streams[5].size - 1001]              # ends inside a byte of a frame
    nch = len(streams)
    n = max(s.size for s in streams)
    buf = np.zeros((nch, n), np.float32)
    for i, s in enumerate(streams):
        buf[i, : s.size] = s
    nb = torch.tensor([s.size for s in streams], dtype=torch.int32, device="cuda")
    whole = fsk.rs41_framesync_device(torch.from_numpy(buf).cuda(), n, nb, nch, ecc=ecc)
    rng = np.random.default_rng(5 + ecc)
    sync = fsk.Rs41FrameSync(nch, ecc=ecc)
    pos = np.zeros(nch, np.int64)
    got, npush = [], 0
    while True:
        rem = np.array([s.size for s in streams]) - pos
        if rem.max() == 0:
            got += sync.push(torch.zeros((nch, 8), dtype=torch.float32, device="cuda"), 8, 0, flush=True)     # empty push that only flushes
            break
        chunk = int(rng.choice([1, 3, 7, 8, 9, 63, 64, 65, 500, 4079, 4080, 4144, 9000, 20000]))
        take = np.minimum(rem, rng.integers(0, chunk + 1, nch))
        m = max(int(take.max()), 1)
        part = np.zeros((nch, m), np.float32)
        for c in range(nch):
            part[c, : take[c]] = streams[c][pos[c]: pos[c] + take[c]]
        got += sync.push(torch.from_numpy(part).cuda(), m, torch.from_numpy(take.astype(np.int32)).cuda())
        pos += take
        npush += 1
    assert npush > 20
    key = lambda g: (g[0], g[2])
    got.sort(key=key)
    whole = sorted(whole, key=key)
    assert len(got) == len(whole) > 40
    for g, e in zip(got, whole):
        assert (g[0], g[1], g[2], g[5], g[6], g[7]) == (e[0], e[1], e[2], e[5], e[6], e[7]), (g[:3], g[5:], e[:3], e[5:])
        assert g[4].tobytes() == e[4].tobytes() and abs(g[3] - e[3]) < 1e-6
