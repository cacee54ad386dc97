"""N > 1 path on CPU: world_size 2, gloo, the simulator build standing in for the device.

Checks the only cross-rank dependency of the path -- the reference's channel-order float32 noise
mean (stationary.py:61-64) chained across ranks -- gives thresholds BIT-EQUAL to a single process
holding all channels, and that the all-gathered waveform equals the single-process result.
"""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SR = 16000
KW = dict(chunk_size=2500, padding=400)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker_interleaved(rank, world, port, result_path):
    sys.path.insert(0, ROOT)
    from noisereduce_b200.device import DeviceGate
    from noisereduce_b200.parallel import gathered_noise_stats
    from tests.cusim_util import cusim_library
    from tests.synth_host import synth_small
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    y = torch.from_numpy(synth_small(C=4, n=6000))
    x_local = y[rank::world].contiguous()                 # global channel c * world + rank
    dg = DeviceGate(sr=SR, stationary=True, lib=cusim_library(), **KW)
    gathered_noise_stats(dg, x_local, rank, world)
    out = dg.run(x_local)
    full = torch.empty((x_local.shape[0], world, x_local.shape[1]))
    for c in range(x_local.shape[0]):
        dist.all_gather_into_tensor(full[c].view(-1), out[c])
    if rank == 0:
        np.savez(result_path, full=full.reshape(-1, x_local.shape[1]).numpy(), thr=dg.gate.noise_threshold())
    dist.destroy_process_group()


def _worker(rank, world, port, result_path):
    sys.path.insert(0, ROOT)
    from noisereduce_b200.device import DeviceGate
    from noisereduce_b200.parallel import sharded_reduce_noise
    from tests.cusim_util import cusim_library
    from tests.synth_host import synth_small
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    y = torch.from_numpy(synth_small(C=4, n=6000))
    cpr = y.shape[0] // world
    x_local = y[rank * cpr: (rank + 1) * cpr].contiguous()
    dg = DeviceGate(sr=SR, stationary=True, lib=cusim_library(), **KW)
    full = sharded_reduce_noise(dg, x_local, rank, world)
    thr = dg.gate.noise_threshold()
    gathered = [None] * world
    dist.all_gather_object(gathered, thr)
    if rank == 0:
        np.savez(result_path, full=full.numpy(), thr0=gathered[0], thr1=gathered[1])
    dist.destroy_process_group()


def test_two_ranks_equal_one_process(tmp_path):
    from noisereduce_b200.device import DeviceGate
    from tests.cusim_util import cusim_library
    from tests.synth_host import synth_small
    cusim_library()                                   # build once, before forking workers
    result = str(tmp_path / "res.npz")
    mp.spawn(_worker, args=(2, _free_port(), result), nprocs=2, join=True)
    r = np.load(result)
    y = torch.from_numpy(synth_small(C=4, n=6000))
    dg = DeviceGate(sr=SR, stationary=True, lib=cusim_library(), **KW)
    dg.noise_stats(y)
    single = dg.run(y).numpy()
    thr = dg.gate.noise_threshold()
    assert np.array_equal(r["thr0"], thr) and np.array_equal(r["thr1"], thr)     # bit-equal thresholds
    assert np.array_equal(r["full"], single)                                    # identical waveform


def test_interleaved_ownership_equals_one_process(tmp_path):
    from noisereduce_b200.device import DeviceGate
    from tests.cusim_util import cusim_library
    from tests.synth_host import synth_small
    cusim_library()
    result = str(tmp_path / "res2.npz")
    mp.spawn(_worker_interleaved, args=(2, _free_port(), result), nprocs=2, join=True)
    r = np.load(result)
    y = torch.from_numpy(synth_small(C=4, n=6000))
    dg = DeviceGate(sr=SR, stationary=True, lib=cusim_library(), **KW)
    dg.noise_stats(y)
    assert np.array_equal(r["thr"], dg.gate.noise_threshold())
    assert np.array_equal(r["full"], dg.run(y).numpy())


def _worker_ring(rank, world, port, result_path):
    sys.path.insert(0, ROOT)
    from noisereduce_b200.device import DeviceGate
    from noisereduce_b200.parallel import chained_noise_stats, sharded_run_slab_ring
    from tests.cusim_util import cusim_library
    from tests.synth_host import synth_small
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    y = torch.from_numpy(synth_small(C=4, n=11000))
    cpr = y.shape[0] // world
    x_local = y[rank * cpr: (rank + 1) * cpr].contiguous()
    dg = DeviceGate(sr=SR, stationary=True, lib=cusim_library(), chunk_size=3000, padding=400)
    chained_noise_stats(dg, x_local, rank, world)
    res = {}
    for slab_chunks in (1, 3):                       # 4 slabs (ragged last) / 2 slabs (ragged last)
        full = torch.zeros((world * cpr, y.shape[1]))
        sums = sharded_run_slab_ring(
            dg, x_local, world, slab_chunks=slab_chunks,
            consume=lambda g, first, si: (full.view(world, cpr, -1)[:, :, first: first + g.shape[2]].copy_(g),
                                          float(g.double().sum()))[1])
        res[f"full{slab_chunks}"] = full.numpy()
        res[f"sums{slab_chunks}"] = np.array(sums)
    allsums = [None] * world
    dist.all_gather_object(allsums, res["sums1"].tolist())
    if rank == 0:
        np.savez(result_path, same_checksums=np.array(allsums[0] == allsums[1]), **res)
    dist.destroy_process_group()


def test_slab_ring_gather_equals_one_process(tmp_path):
    """Config 5's data path at world_size 2: slabs of the chunk grid are denoised densely (set_range +
    virtual output base), all-gathered into ring slots and consumed; the stitched result equals one process."""
    from noisereduce_b200.device import DeviceGate
    from tests.cusim_util import cusim_library
    from tests.synth_host import synth_small
    cusim_library()
    result = str(tmp_path / "res3.npz")
    mp.spawn(_worker_ring, args=(2, _free_port(), result), nprocs=2, join=True)
    r = np.load(result)
    y = torch.from_numpy(synth_small(C=4, n=11000))
    dg = DeviceGate(sr=SR, stationary=True, lib=cusim_library(), chunk_size=3000, padding=400)
    dg.noise_stats(y)
    single = dg.run(y).numpy()
    assert np.array_equal(r["full1"], single) and np.array_equal(r["full3"], single)
    assert bool(r["same_checksums"]) and len(r["sums1"]) == 4 and len(r["sums3"]) == 2
    assert abs(r["sums1"].sum() - single.astype(np.float64).sum()) < 1e-6
