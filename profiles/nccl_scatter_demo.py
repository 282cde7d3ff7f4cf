#!/usr/bin/env python
"""Multi-GPU evidence for SURVEY §8e: rank 0 holds a shared capture, NCCL scatters the channel blocks over NVLink,
every rank demodulates its shard on its own GPU, per-channel frame counts are gathered and rank 0 checks a few
channels of EVERY shard bit-for-bit against the oracle.
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 profiles/nccl_scatter_demo.py"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
from radiosonde_auto_rx_b200 import fsk, shard, siggen

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
nchan, nsamp = 64 * world, 1 << 17
iq = siggen.fsk_batch(7000, nchan, nsamp) if rank == 0 else None
torch.cuda.synchronize(); dist.barrier()
t0 = time.perf_counter()
mine = shard.scatter_capture(iq, nchan, nsamp, src=0, device=f"cuda:{local}")
torch.cuda.synchronize(); t_sc = time.perf_counter() - t0
lo, hi = shard.shard_range(nchan, world, rank)
b = fsk.FskBatch(48000, 4800, 2, 5, 300, 1000, 5000, -5000, 5000, 1, nchan=hi - lo, device=local)
maxf = b.max_frames(nsamp)
out = torch.zeros((hi - lo, maxf * b.Nbits), dtype=torch.float32, device="cuda")
b.demod_device(mine.contiguous(), "cs16", nsamp, out, soft=True, in_stride=nsamp)
torch.cuda.synchronize()
frames, consumed = b.counts()
allf = shard.gather_counts(frames, nchan, device=f"cuda:{local}")
# every rank checks two of its own channels against the oracle
from oracle import pyoracle as po
mine_h = mine.cpu().numpy(); o = out.cpu().numpy(); bad = 0
for c in (0, hi - lo - 1):
    ref, _, recs = po.Oracle(po.RS41).run_stream(mine_h[c])
    bad += not (frames[c] == len(recs) and o[c, : frames[c] * b.Nbits].tobytes() == ref.tobytes())
badt = torch.tensor([bad], device="cuda"); dist.all_reduce(badt)
if rank == 0:
    print(json.dumps({"world": world, "channels": nchan, "nsamp": nsamp, "scatter_ms": t_sc * 1e3,
                      "scatter_GBs": nchan * nsamp * 4 * (world - 1) / world / t_sc / 1e9,
                      "frames_total": int(allf.sum()), "oracle_mismatches": int(badt.item()), "backend": "nccl"}))
dist.destroy_process_group()
