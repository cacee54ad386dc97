"""Multi-GPU: one process per GPU, channels sharded across ranks.

Every (chunk, channel) unit is independent, so the data path needs no collective.  The only
cross-rank dependency is the stationary noise threshold, which the reference computes from the mean
over ALL channels in channel order and in the input dtype (stationary.py:61-64).  To reproduce that
float32 sum bit for bit, ranks chain it: rank r continues the running sum of rank r-1 with its own
channels (one 2.4 MB point-to-point hop per rank), the last rank divides, and the collapsed noise clip
is broadcast.  The result is gathered with ONE all_gather of the final waveform (north_star).
torch.distributed (NCCL on GPUs, gloo in the CPU tests) is the plumbing.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def chained_noise_stats(dg, x_local: torch.Tensor, rank: int, world: int, group=None):
    """Exact multi-rank version of SpectralGateStationary.__init__'s noise statistics."""
    p = dg.gate.params
    C, N = x_local.shape
    n = N
    if p.clip_noise and p.chunk_size > 0 and n > p.chunk_size:
        n = int(p.chunk_size)
    acc = torch.zeros(n, dtype=torch.float32, device=x_local.device)
    if rank > 0:
        dist.recv(acc, src=rank - 1, group=group)
    stream = torch.cuda.current_stream().cuda_stream if x_local.is_cuda else None
    dg.gate.channel_sum_device(x_local.data_ptr(), np.float32, C, n, x_local.stride(0), acc.data_ptr(),
                               init=(rank == 0), stream=stream)
    if x_local.is_cuda:
        torch.cuda.current_stream().synchronize()
    if rank < world - 1:
        dist.send(acc, dst=rank + 1, group=group)
    mean = acc / np.float32(world * C) if rank == world - 1 else acc      # numpy: sum / count in float32
    dist.broadcast(mean, src=world - 1, group=group)
    dg.gate.noise_stats_collapsed_device(mean.data_ptr(), np.float32, n, stream=stream)
    return mean


def sharded_reduce_noise(dg, x_local: torch.Tensor, rank: int, world: int, gather=True, group=None):
    """Denoise this rank's channel shard; optionally all-gather the [world*C, N] result."""
    if dg.stationary:
        chained_noise_stats(dg, x_local, rank, world, group)
    y_local = dg.run(x_local)
    if not gather or world == 1:
        return y_local
    full = torch.empty((world * x_local.shape[0], x_local.shape[1]), dtype=x_local.dtype, device=x_local.device)
    dist.all_gather_into_tensor(full, y_local, group=group)
    return full


def sharded_run_overlapped(dg, x_local: torch.Tensor, out_local: torch.Tensor, gathered: torch.Tensor, world: int,
                           comm_stream, groups: int = 8, group=None):
    """Denoise this rank's channels group by group and all-gather each finished group on `comm_stream`
    while the next group is being computed.  `gathered` is the final [world*C, N] waveform in channel
    order; each group's all-gather writes rank r's rows straight into gathered[r*C + g0 : r*C + g1].
    Thresholds must already be set (chained_noise_stats)."""
    C, N = x_local.shape
    g3 = gathered.view(world, C, N)
    gs = (C + groups - 1) // groups
    cur = torch.cuda.current_stream()
    for g0 in range(0, C, gs):
        g1 = min(C, g0 + gs)
        dg.run(x_local[g0:g1], out_local[g0:g1])
        ev = torch.cuda.Event()
        ev.record(cur)
        comm_stream.wait_event(ev)
        with torch.cuda.stream(comm_stream):
            dist.all_gather([g3[r, g0:g1] for r in range(world)], out_local[g0:g1], group=group)
    cur.wait_stream(comm_stream)
    return gathered


# ---- interleaved channel ownership: zero-copy, overlappable gathers --------------------------------
def gathered_noise_stats(dg, x_local: torch.Tensor, rank: int, world: int, group=None):
    """Ownership: global channel c * world + rank.  Every rank gathers the (small) noise clip of all
    channels, orders it globally and runs the ordinary single-process noise statistics on it, so the
    thresholds are exactly those of one process holding all channels (stationary.py:61-81)."""
    p = dg.gate.params
    C, N = x_local.shape
    n = N
    if p.clip_noise and p.chunk_size > 0 and n > p.chunk_size:
        n = int(p.chunk_size)
    clip = x_local[:, :n].contiguous()
    allc = torch.empty((world, C, n), dtype=clip.dtype, device=clip.device)
    dist.all_gather_into_tensor(allc.view(-1), clip.view(-1), group=group)
    ordered = allc.permute(1, 0, 2).reshape(C * world, n).contiguous()      # row c*world + r
    dg.noise_stats(ordered)
    return ordered


def interleaved_run_overlapped(dg, x_local: torch.Tensor, out_local: torch.Tensor, gathered: torch.Tensor, world: int,
                               comm_stream, groups: int = 8, group=None):
    """gathered: [C, world, N] == the final [C*world, N] waveform with global channel c*world + rank.
    Each finished local channel is all-gathered straight into gathered[c] (contiguous, zero-copy) on
    `comm_stream` while the next channel group is being computed."""
    C, N = x_local.shape
    gs = (C + groups - 1) // groups
    cur = torch.cuda.current_stream()
    for g0 in range(0, C, gs):
        g1 = min(C, g0 + gs)
        dg.run(x_local[g0:g1], out_local[g0:g1])
        ev = torch.cuda.Event()
        ev.record(cur)
        comm_stream.wait_event(ev)
        with torch.cuda.stream(comm_stream):
            for c in range(g0, g1):
                dist.all_gather_into_tensor(gathered[c].view(-1), out_local[c], group=group)
    cur.wait_stream(comm_stream)
    return gathered
