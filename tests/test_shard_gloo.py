"""N>1 host logic on CPU: world_size-2 gloo processes scatter a shared capture by channel, each rank demodulates its
shard (here with the oracle standing in for the GPU), counters are gathered and must equal the single-process run."""
import os
import socket

import numpy as np
import pytest

from radiosonde_auto_rx_b200 import shard


def test_shard_range_partitions_exactly():
    for n in (1, 7, 256, 1024, 1025):
        for w in (1, 2, 3, 8):
            rs = [shard.shard_range(n, w, r) for r in range(w)]
            assert rs[0][0] == 0 and rs[-1][1] == n
            assert all(rs[i][1] == rs[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in rs]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, nchan, nsamp, q):
    import torch.distributed as dist
    from oracle import pyoracle as po
    from radiosonde_auto_rx_b200 import siggen
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    iq = siggen.fsk_batch(4242, nchan, nsamp, workers=2) if rank == 0 else None
    mine = shard.scatter_capture(iq, nchan, nsamp, src=0).numpy()
    lo, hi = shard.shard_range(nchan, world, rank)
    assert mine.shape[0] == hi - lo
    frames, sums = [], []
    for c in range(hi - lo):
        out, cons, recs = po.Oracle(po.RS41).run_stream(mine[c])
        frames.append(len(recs))
        sums.append(int(np.frombuffer(out.tobytes(), dtype=np.uint32).astype(np.uint64).sum() % (1 << 62)))
    f_all = shard.gather_counts(frames, nchan)
    s_all = shard.gather_counts(sums, nchan)
    if rank == 0:
        q.put((f_all.tolist(), s_all.tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_scatter_shard_gather_world2(built):
    import torch.multiprocessing as mp
    from oracle import pyoracle as po
    from radiosonde_auto_rx_b200 import siggen
    nchan, nsamp, world = 5, 20_000, 2
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, world, port, nchan, nsamp, q)) for r in range(world)]
    [p.start() for p in ps]
    f_all, s_all = q.get(timeout=120)
    [p.join(60) for p in ps]
    assert all(p.exitcode == 0 for p in ps)
    iq = siggen.fsk_batch(4242, nchan, nsamp, workers=2)
    for c in range(nchan):
        out, _, recs = po.Oracle(po.RS41).run_stream(iq[c])
        assert f_all[c] == len(recs)
        assert s_all[c] == int(np.frombuffer(out.tobytes(), dtype=np.uint32).astype(np.uint64).sum() % (1 << 62))


def _wb_worker(rank, world, port, q):
    import torch.distributed as dist
    from oracle import pyoracle as po
    from radiosonde_auto_rx_b200 import siggen
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sr, hz = 1000000, [200e3, -300e3, 50e3]
    n = int(sr * 1.3)
    wide = siggen.wideband_mix(12, n, sr, [("RS41", 200e3), ("LMS6", -300e3)], snr_db=8.0) if rank == 0 else None
    mine = shard.broadcast_capture(wide, n, src=0).numpy()
    lo, hi = shard.shard_range(len(hz), world, rank)                 # each rank owns a block of candidate frequencies
    codes = [po.dft_oracle_run_wideband(mine, sr, hz[c] / sr, 15.0, True, False)[0] for c in range(lo, hi)]   # the oracle stands in for the GPU
    allc = shard.gather_counts(codes, len(hz))
    if rank == 0:
        q.put((allc.tolist(), int(np.abs(mine.astype(np.int64)).sum())))
    dist.barrier()
    dist.destroy_process_group()


def test_broadcast_capture_shard_candidates_world2(built):
    """wideband mode: one capture to every rank, candidates sharded, exit codes gathered == single-process run"""
    import torch.multiprocessing as mp
    from oracle import pyoracle as po
    from radiosonde_auto_rx_b200 import siggen
    world = 2
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_wb_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in ps]
    codes, chk = q.get(timeout=180)
    [p.join(60) for p in ps]
    assert all(p.exitcode == 0 for p in ps)
    sr = 1000000
    wide = siggen.wideband_mix(12, int(sr * 1.3), sr, [("RS41", 200e3), ("LMS6", -300e3)], snr_db=8.0)
    assert chk == int(np.abs(wide.astype(np.int64)).sum())
    exp = [po.dft_oracle_run_wideband(wide, sr, f / sr, 15.0, True, False)[0] for f in (200e3, -300e3, 50e3)]
    assert codes == exp and exp[0] == 3 and exp[1] == 8 and exp[2] == 0


def _bs_worker(rank, world, port, C, B, T, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    full = np.random.default_rng(99).integers(-32768, 32767, (world * C, T * B, 2), dtype=np.int16)   # every rank can rebuild the expectation
    cap = torch.from_numpy(full) if rank == 0 else None
    bs = shard.BlockScatter(C, B, "cpu", src=0)
    ok = True
    for step in range(2):                                   # the two receive buffers are reused across steps
        got, order = [], []
        bs.step(cap, T, lambda buf, t, stream: (got.append(buf.clone().numpy()), order.append(t), stream is None))
        ok = ok and order == list(range(T))
        ok = ok and np.array_equal(np.concatenate(got, axis=1), full[rank * C:(rank + 1) * C])
    q.put((rank, bool(ok)))
    dist.barrier()
    dist.destroy_process_group()


def test_block_scatter_pipeline_world2():
    """The streaming scatter bench.py --scatter times on GPUs (block t+1 in flight while block t is consumed), on the gloo backend:
    every rank must see its channel block of every time block, in order, through the two alternating buffers."""
    import torch.multiprocessing as mp
    world, C, B, T = 2, 3, 500, 5
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_bs_worker, args=(r, world, port, C, B, T, q)) for r in range(world)]
    [p.start() for p in ps]
    res = dict(q.get(timeout=120) for _ in range(world))
    [p.join(60) for p in ps]
    assert all(p.exitcode == 0 for p in ps)
    assert res == {0: True, 1: True}


def test_block_scatter_single_rank_is_a_copy():
    import torch
    C, B, T = 2, 64, 3
    full = np.random.default_rng(5).integers(-100, 100, (C, T * B, 2), dtype=np.int16)
    bs = shard.BlockScatter(C, B, "cpu")
    got = []
    bs.step(torch.from_numpy(full), T, lambda buf, t, stream: got.append(buf.clone().numpy()))
    assert np.array_equal(np.concatenate(got, axis=1), full)
