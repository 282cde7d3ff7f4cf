"""Multi-GPU plumbing for the channel-sharded path.

The hot path partitions by channel: GPU g owns a contiguous block of narrowband IQ streams and their persistent
demodulator state for ever; no data-path collective exists.  A collective is used only to DISTRIBUTE input when one
rank holds a shared capture: ``scatter_capture`` (torch.distributed.scatter -- NCCL over NVLink on GPUs, gloo on CPU
for the tests) hands every rank its [channels_g, nsamp, 2] slab; results (bits, detections) stay per rank and only
tiny per-channel counters are gathered.  For a WIDEBAND capture (BASELINE cfg5) every rank needs the whole
buffer and owns a block of candidate frequencies instead: ``broadcast_capture`` (one NCCL broadcast, 4 B per wide sample)
followed by ``shard_range`` over the candidate list.  ``BlockScatter`` is the streaming form bench.py --scatter times on GPUs: the capture
is scattered one time block after the other and block t+1 travels while block t is demodulated.
"""
from __future__ import annotations

import numpy as np


def shard_range(nchan_total: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous channel block [lo, hi) owned by `rank`; sizes differ by at most one."""
    base, rem = divmod(nchan_total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def scatter_capture(iq_root, nchan_total: int, nsamp: int, src: int = 0, device=None):
    """Rank `src` passes the full capture (int16 [nchan_total, nsamp, 2], torch tensor or numpy); every rank gets its
    shard as a torch tensor on `device`.  Shards are padded to equal size for the collective and trimmed after."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(), dist.get_rank()
    per = -(-nchan_total // world)
    lo, hi = shard_range(nchan_total, world, rank)
    dev = device if device is not None else "cpu"
    # collectives move raw bytes (gloo has no int16): view the slabs as uint8
    recv = torch.empty((per, nsamp * 4), dtype=torch.uint8, device=dev)
    chunks = None
    if rank == src:
        full = torch.as_tensor(np.ascontiguousarray(iq_root) if isinstance(iq_root, np.ndarray) else iq_root).to(dev)
        chunks = []
        for r in range(world):
            a, b = shard_range(nchan_total, world, r)
            t = torch.zeros((per, nsamp, 2), dtype=torch.int16, device=dev)
            t[: b - a] = full[a:b]
            chunks.append(t.view(torch.uint8).reshape(per, nsamp * 4))
    dist.scatter(recv, chunks, src=src)
    return recv.view(torch.int16).reshape(per, nsamp, 2)[: hi - lo]


def broadcast_capture(wide_root, nwide: int, src: int = 0, device=None):
    """Rank `src` passes one wideband capture (int16 [nwide, 2]); every rank returns it as a torch tensor on `device`.
    (torch.distributed.broadcast: NCCL over NVLink / NVSwitch on GPUs -- 38 MB per second of 9.6 MHz capture, gloo in the CPU test.)"""
    import torch
    import torch.distributed as dist
    dev = device if device is not None else "cpu"
    buf = torch.empty(nwide * 4, dtype=torch.uint8, device=dev)              # raw bytes: gloo has no int16
    if dist.get_rank() == src:
        full = torch.as_tensor(np.ascontiguousarray(wide_root) if isinstance(wide_root, np.ndarray) else wide_root)
        buf.copy_(full.reshape(-1).view(torch.uint8))
    dist.broadcast(buf, src=src)
    return buf.view(torch.int16).reshape(nwide, 2)


def gather_counts(local_counts, nchan_total: int, dst: int = 0, device=None):
    """Per-channel int counters of every rank -> one [nchan_total] array on `dst` (None elsewhere)."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(), dist.get_rank()
    per = -(-nchan_total // world)
    dev = device if device is not None else "cpu"
    t = torch.zeros(per, dtype=torch.int64, device=dev)
    t[: len(local_counts)] = torch.as_tensor(np.asarray(local_counts, dtype=np.int64)).to(dev)
    out = [torch.zeros_like(t) for _ in range(world)] if rank == dst else None
    dist.gather(t, out, dst=dst)
    if rank != dst:
        return None
    res = []
    for r in range(world):
        a, b = shard_range(nchan_total, world, r)
        res.append(out[r][: b - a].cpu().numpy())
    return np.concatenate(res)


class BlockScatter:
    """Time-blocked scatter of a capture that one rank holds: [world * C channels, T * B samples, 2] int16 on rank `src`'s device.
    Every rank receives its C channels one block of B samples after the other into two alternating buffers; block t+1 is in flight
    while `consume` works on block t (profiles/r02_notes.md section 6: 1 / 2 / 4 GPUs).  On CUDA devices the scatter is issued on its own
    stream (NCCL over NVLink; NCCL has no 16-bit integer type, so one cs16 sample (I, Q) travels as one int32) and `consume` must enqueue
    its work on `comp_stream`; on CPU tensors (gloo, the tests) the same schedule runs synchronously.  Without a process group (one
    rank) the "scatter" is a device copy, so the single-GPU figure measures the same pipeline."""

    def __init__(self, channels_per_rank: int, block: int, device, src: int = 0):
        import torch
        import torch.distributed as dist
        self.C, self.B, self.src = int(channels_per_rank), int(block), int(src)
        self.dist = dist if dist.is_available() and dist.is_initialized() else None
        self.world = self.dist.get_world_size() if self.dist else 1
        self.rank = self.dist.get_rank() if self.dist else 0
        self.device = torch.device(device)
        self.cuda = self.device.type == "cuda"
        self.recv = [torch.empty((self.C, self.B, 2), dtype=torch.int16, device=self.device) for _ in range(2)]
        if self.cuda:
            self.copy_stream, self.comp_stream = torch.cuda.Stream(self.device), torch.cuda.Stream(self.device)
            self.ev_rx = [torch.cuda.Event() for _ in range(2)]
            self.ev_done = [torch.cuda.Event() for _ in range(2)]
        else:
            self.copy_stream = self.comp_stream = None

    def _scatter_block(self, cap, t: int) -> None:
        import torch
        C, B, buf = self.C, self.B, self.recv[t % 2]
        if self.world > 1:
            lst = None
            if self.rank == self.src:
                lst = [cap[r * C:(r + 1) * C, t * B:(t + 1) * B].contiguous().view(torch.int32) for r in range(self.world)]
            self.dist.scatter(buf.view(torch.int32), lst, src=self.src)
        else:
            buf.copy_(cap[:, t * B:(t + 1) * B])

    def _post(self, cap, t: int) -> None:
        """block t -> recv[t % 2] (on the copy stream, after the consumer that last read that buffer)"""
        import torch
        if not self.cuda:
            self._scatter_block(cap, t)
            return
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self.ev_done[t % 2])
            self._scatter_block(cap, t)
            self.ev_rx[t % 2].record(self.copy_stream)

    def step(self, cap, nblocks: int, consume) -> None:
        """One pass over the capture: `consume(buffer [C, B, 2] int16, t, comp_stream)` is called for t = 0 .. nblocks-1 in order."""
        self._post(cap, 0)
        for t in range(nblocks):
            if t + 1 < nblocks:
                self._post(cap, t + 1)
            if self.cuda:
                self.comp_stream.wait_event(self.ev_rx[t % 2])
            consume(self.recv[t % 2], t, self.comp_stream)
            if self.cuda:
                self.ev_done[t % 2].record(self.comp_stream)
