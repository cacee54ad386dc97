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
    # the reference's channel mean runs in the input dtype for float32 and in float64 for int16 / float64
    # (stationary.py:61-64 through numpy's promotion); the running sum follows the same rule
    if x_local.dtype not in (torch.float32, torch.int16, torch.float64):
        raise TypeError("sharded noise statistics take float32 / int16 / float64 shards")
    acc_t = torch.float32 if x_local.dtype == torch.float32 else torch.float64
    np_in = {torch.float32: np.float32, torch.int16: np.int16, torch.float64: np.float64}[x_local.dtype]
    np_acc = np.float32 if acc_t == torch.float32 else np.float64
    acc = torch.zeros(n, dtype=acc_t, device=x_local.device)
    if rank > 0:
        dist.recv(acc, src=rank - 1, group=group)
    stream = torch.cuda.current_stream().cuda_stream if x_local.is_cuda else None
    dg.gate.channel_sum_device(x_local.data_ptr(), np_in, C, n, x_local.stride(0), acc.data_ptr(),
                               init=(rank == 0), stream=stream)
    if x_local.is_cuda:
        torch.cuda.current_stream().synchronize()
    if rank < world - 1:
        dist.send(acc, dst=rank + 1, group=group)
    mean = acc / np_acc(world * C) if rank == world - 1 else acc      # numpy: sum / count in the accumulation dtype
    dist.broadcast(mean, src=world - 1, group=group)
    dg.gate.noise_stats_collapsed_device(mean.data_ptr(), np_acc, n, stream=stream)
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


# ---- config 5: the gathered result does not fit one GPU -> gather slab by slab into a ring ------------------
def make_slab_ring(C: int, slab_len: int, world: int, dtype, device):
    """Two dense [C, slab] compute buffers and two [world, C, slab] gather slots."""
    bufs = [torch.empty((C, slab_len), dtype=dtype, device=device) for _ in range(2)]
    ring = [torch.empty((world, C, slab_len), dtype=dtype, device=device) for _ in range(2)] if world > 1 else None
    return bufs, ring


def sharded_run_slab_ring(dg, x_local: torch.Tensor, world: int, comm_stream=None, slab_chunks: int = 1,
                          consume=None, group=None, buffers=None, peer: "PeerGather | None" = None):
    """Channel-sharded run whose all-gathered result (world * C * N samples) is larger than one GPU's memory
    (BASELINE config 5: 512 ch x 60 min = 354 GB).  The recording is walked in slabs of `slab_chunks` chunks
    of the reference's chunk grid: each slab is denoised into one of two dense [C, slab] buffers
    (DeviceGate.run_chunks), all-gathered on `comm_stream` into one of two [world, C, slab] ring slots while
    the next slab is being computed, and handed to `consume(gathered_view, first_sample, slab_index)` (run on
    the communication stream: a checksum, a writer, the next stage); nothing else is retained.
    Returns the list of consume() results.  Thresholds must already be set (chained_noise_stats).
    With `peer` (a PeerGather of shape [2, world, C, slab]) the ring lives in symmetric memory: the kernels
    write slab s into peer.buf[s & 1, rank] and the copy engines push it to every peer (no NCCL kernels)."""
    if peer is not None:
        return _slab_ring_peer(dg, x_local, world, slab_chunks, consume, peer)
    C, N = x_local.shape
    cs = int(dg.gate.params.chunk_size)
    n_chunks = (N - 1) // cs + 1
    slab_len = slab_chunks * cs
    bufs, ring = buffers if buffers is not None else make_slab_ring(C, slab_len, world, x_local.dtype, x_local.device)
    cuda = x_local.is_cuda
    cur = torch.cuda.current_stream() if cuda else None
    gathered_free = [None, None]           # events: ring slot / dense buffer consumed
    results = []
    for si, first in enumerate(range(0, n_chunks, slab_chunks)):
        last = min(n_chunks - 1, first + slab_chunks - 1)
        k = si & 1
        if cuda and gathered_free[k] is not None:
            cur.wait_event(gathered_free[k])                   # slot k's previous gather + consume finished
        view = dg.run_chunks(x_local, bufs[k], first, last)    # [C, n_s]
        n_s = view.shape[1]
        if cuda:
            ev = torch.cuda.Event()
            ev.record(cur)
            comm_stream.wait_event(ev)
        ctx = torch.cuda.stream(comm_stream) if cuda else _Null()
        with ctx:
            if world > 1:
                if n_s == slab_len:
                    dist.all_gather_into_tensor(ring[k].view(-1), bufs[k].view(-1), group=group)
                    g = ring[k]
                else:                                          # ragged last slab: gather the dense part only
                    part = view.contiguous()
                    g = ring[k].view(-1)[: world * C * n_s].view(world, C, n_s)
                    dist.all_gather_into_tensor(g.view(-1), part.view(-1), group=group)
            else:
                g = view.unsqueeze(0)
            results.append(consume(g, first * cs, si) if consume is not None else None)
            if cuda:
                gathered_free[k] = torch.cuda.Event()
                gathered_free[k].record(comm_stream)
    if cuda:
        cur.wait_stream(comm_stream)
    return results


class _Null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


# ---- peer-memory gather: results pushed into every peer's buffer by the copy engines over NVLink ----------
class PeerGather:
    """The all-gather of the final waveform without collective kernels.

    The gathered [world, C, N] tensor lives in symmetric memory (torch.distributed._symmetric_memory: every
    rank's buffer is mapped into every peer's address space over NVLink / NVSwitch).  A rank's kernels write
    its own rows straight into its slice of the local buffer; as soon as a channel group is finished, plain
    device-to-device copies -- copy engines, no SMs, one stream per peer -- push those rows into the same
    slice of every peer's buffer while the next group is computed.  A device-side barrier at the end of the
    step makes all pushes visible.  Compared with NCCL all-gather kernels running next to the persistent
    gate kernels this leaves all SMs to the gate (no reserve_sms) and removes the kernels' HBM contention.
    """

    def __init__(self, world: int, rank: int, shape, dtype, device, group=None, splits: int = 1):
        import torch.distributed._symmetric_memory as symm
        self.world, self.rank, self.shape = world, rank, tuple(shape)
        self.splits = max(1, int(splits))            # concurrent sub-copies per peer (more copy engines in flight)
        self.buf = symm.empty(self.shape, dtype=dtype, device=device)
        self.h = symm.rendezvous(self.buf, group if group is not None else dist.group.WORLD)
        self.views = [self.buf if r == rank else self.h.get_buffer(r, self.shape, dtype) for r in range(world)]
        self.streams = [torch.cuda.Stream(device=device) for _ in range(world)]
        self.sub = [[torch.cuda.Stream(device=device) for _ in range(self.splits - 1)] for _ in range(world)]
        self.h.barrier()

    def push(self, index, after: "torch.cuda.Event"):
        """Copy self.buf[index] into the same place of every peer's buffer once `after` has fired."""
        src = self.buf[index]
        for r in range(self.world):
            if r == self.rank:
                continue
            dst = self.views[r][index]
            rows = src.shape[0]
            parts = min(self.splits, rows) if src.dim() >= 2 else 1
            step = (rows + parts - 1) // parts if parts > 1 else rows
            for i in range(parts):
                st = self.streams[r] if i == 0 else self.sub[r][i - 1]
                st.wait_event(after)
                with torch.cuda.stream(st):
                    if parts == 1:
                        dst.copy_(src, non_blocking=True)
                    else:
                        dst[i * step: (i + 1) * step].copy_(src[i * step: (i + 1) * step], non_blocking=True)

    def finish(self):
        """Join the push streams into the current stream, then a device-side barrier across ranks."""
        cur = torch.cuda.current_stream()
        for r in range(self.world):
            if r != self.rank:
                cur.wait_stream(self.streams[r])
                for st in self.sub[r]:
                    cur.wait_stream(st)
        self.h.barrier()


def sharded_run_peer_push(dg, x_local: torch.Tensor, pg: PeerGather, groups: int = 8):
    """Channel-sharded run with the peer-memory gather: pg.buf is the final [world, C, N] waveform on every
    rank when this returns (stream-ordered).  Thresholds must already be set (chained_noise_stats)."""
    C, N = x_local.shape
    gs = (C + groups - 1) // groups
    cur = torch.cuda.current_stream()
    mine = pg.buf[pg.rank]
    for g0 in range(0, C, gs):
        g1 = min(C, g0 + gs)
        dg.run(x_local[g0:g1], mine[g0:g1])
        ev = torch.cuda.Event()
        ev.record(cur)
        pg.push((pg.rank, slice(g0, g1)), ev)
    pg.finish()
    return pg.buf


def _slab_ring_peer(dg, x_local, world, slab_chunks, consume, pg: PeerGather):
    C, N = x_local.shape
    cs = int(dg.gate.params.chunk_size)
    n_chunks = (N - 1) // cs + 1
    cur = torch.cuda.current_stream()
    comm = pg.streams[pg.rank]                       # consume() runs here
    slot_free = [None, None]
    results = []
    for si, first in enumerate(range(0, n_chunks, slab_chunks)):
        last = min(n_chunks - 1, first + slab_chunks - 1)
        k = si & 1
        if slot_free[k] is not None:
            cur.wait_event(slot_free[k])             # every rank consumed slot k's previous slab (barrier-ordered)
        view = dg.run_chunks(x_local, pg.buf[k, pg.rank], first, last)
        n_s = view.shape[1]
        ev = torch.cuda.Event()
        ev.record(cur)
        pg.push((k, pg.rank), ev)                    # whole slot rows: dense [C, slab] block, one copy per peer
        comm.wait_event(ev)
        for r in range(world):
            if r != pg.rank:
                comm.wait_stream(pg.streams[r])
                for st in pg.sub[r]:
                    comm.wait_stream(st)
        with torch.cuda.stream(comm):
            pg.h.barrier()                           # all ranks' pushes of this slab have landed
            results.append(consume(pg.buf[k, :, :, :n_s], first * cs, si) if consume is not None else None)
            pg.h.barrier(channel=1)                  # ... and have been consumed everywhere before the slot is reused
            slot_free[k] = torch.cuda.Event()
            slot_free[k].record(comm)
    cur.wait_stream(comm)
    return results


# ---- kernel-issued NVLink stores: b200gate_run_sharded (include/b200gate.h, csrc/gate_peer.cuh) ---------------------
class PeerStore:
    """Symmetric-memory state of the fused gather: the gathered [world, C, N] result, a flag array for the device-side
    epoch barrier, every peer's mapping of both, and the communication stream.  The library does the rest
    (b200gate_run_sharded): gate kernels write this rank's rows into its slice of the local buffer, k_peer_push stores
    every finished channel group into the peers' buffers over NVLink from a few SMs the gate leaves free
    (DeviceGate(reserve_sms=...)), k_peer_barrier ends the step.  No NCCL kernels, no copy engines, no host syncs."""

    def __init__(self, world: int, rank: int, shape, dtype, device, group=None):
        import torch.distributed._symmetric_memory as symm
        grp = group if group is not None else dist.group.WORLD
        self.world, self.rank, self.shape = world, rank, tuple(shape)
        self.buf = symm.empty(self.shape, dtype=dtype, device=device)
        self.h = symm.rendezvous(self.buf, grp)
        self.flags = symm.empty((64,), dtype=torch.int32, device=device)
        self.flags.zero_()
        self.hf = symm.rendezvous(self.flags, grp)
        self._views = [self.buf if r == rank else self.h.get_buffer(r, self.shape, dtype) for r in range(world)]
        self._fviews = [self.flags if r == rank else self.hf.get_buffer(r, (64,), torch.int32) for r in range(world)]
        self.peer_ptrs = [v.data_ptr() for v in self._views]
        self.flag_ptrs = [v.data_ptr() for v in self._fviews]
        self.comm = torch.cuda.Stream(device=device, priority=-1)
        self.epoch = 0
        torch.cuda.synchronize()
        self.hf.barrier()                                   # every rank's flags are zero before the first epoch

    def next_epoch(self) -> int:
        self.epoch += 1
        return self.epoch


def sharded_run_peer_store(dg, x_local: torch.Tensor, ps: PeerStore, groups: int = 8, push_ctas: int = 0):
    """Channel-sharded step with the kernel-issued gather: ps.buf is the final [world, C, N] waveform on every rank
    once the current stream reaches this point.  Thresholds must already be set (chained_noise_stats)."""
    C, N = x_local.shape
    cur = torch.cuda.current_stream()
    dg.gate.run_sharded(x_local.data_ptr(), DeviceGateDtypes[x_local.dtype], C, N, x_local.stride(0), ps.buf.data_ptr(),
                        ps.peer_ptrs, ps.flags.data_ptr(), ps.flag_ptrs, ps.next_epoch(), ps.rank, ps.world, groups,
                        push_ctas, cur.cuda_stream, ps.comm.cuda_stream)
    return ps.buf


DeviceGateDtypes = {torch.float32: np.float32, torch.int16: np.int16, torch.float64: np.float64}


def slab_ring_peer_store(dg, x_local, world, slab_chunks, consume, ps: PeerStore, push_ctas: int = 0):
    """Config-5 slab ring on the kernel-issued gather: ps.buf is [2, world, C, slab]; slab s is denoised into
    ps.buf[s & 1, rank], pushed into every peer's copy by k_peer_push behind the next slab's kernels, published by the
    epoch barrier, consumed, and released by a second barrier before its slot is reused."""
    from . import _cabi
    C, N = x_local.shape
    cs = int(dg.gate.params.chunk_size)
    n_chunks = (N - 1) // cs + 1
    cur = torch.cuda.current_stream()
    comm = ps.comm
    lib = dg.gate.lib
    es = x_local.element_size()
    slab_len = ps.buf.shape[-1]
    slot_free = [None, None]
    results = []
    peers = [r for r in range(world) if r != ps.rank]
    for si, first in enumerate(range(0, n_chunks, slab_chunks)):
        last = min(n_chunks - 1, first + slab_chunks - 1)
        k = si & 1
        if slot_free[k] is not None:
            cur.wait_event(slot_free[k])
        mine = ps.buf[k, ps.rank]
        view = dg.run_chunks(x_local, mine, first, last)
        n_s = view.shape[1]
        ev = torch.cuda.Event()
        ev.record(cur)
        comm.wait_event(ev)
        off = ((k * world + ps.rank) * C) * slab_len * es
        if peers:
            _cabi.peer_push(lib, mine.data_ptr(), [ps.peer_ptrs[r] + off for r in peers], C, n_s * es, slab_len * es,
                            slab_len * es, push_ctas, comm.cuda_stream)
            _cabi.peer_barrier(lib, ps.flags.data_ptr(), ps.flag_ptrs, ps.rank, world, ps.next_epoch(), comm.cuda_stream)
        with torch.cuda.stream(comm):
            results.append(consume(ps.buf[k, :, :, :n_s], first * cs, si) if consume is not None else None)
        if peers:
            _cabi.peer_barrier(lib, ps.flags.data_ptr(), ps.flag_ptrs, ps.rank, world, ps.next_epoch(), comm.cuda_stream)
        slot_free[k] = torch.cuda.Event()
        slot_free[k].record(comm)
    cur.wait_stream(comm)
    return results
