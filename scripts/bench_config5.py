#!/usr/bin/env python
"""BASELINE.json configs[4] (config 5): 512 ch x 60 min @ 48 kHz, stationary, chunk_size = 60 s, channel-sharded
64 ch per GPU, the all-gathered result (354 GB at 8 GPUs) consumed slab by slab through a two-slot ring
(noisereduce_b200/parallel.py: sharded_run_slab_ring).  One JSON line like bench.py's.

    python scripts/bench_config5.py [--minutes 60] [--steps 2] [--warmup 1]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/bench_config5.py ...

`value` counts the samples of ALL ranks' shards (weak scaling: 64 channels per GPU).
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import synth_device, SR, C_PER_GPU, METRIC, UNIT  # noqa: E402
from noisereduce_b200.device import DeviceGate  # noqa: E402
from noisereduce_b200.parallel import chained_noise_stats, make_slab_ring, sharded_run_slab_ring  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--minutes", type=float, default=60.0)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--slab-chunks", type=int, default=1)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    C, n, cs = C_PER_GPU, int(args.minutes * 60 * SR), 60 * SR
    # the 44 GB shard is synthesised 10 minutes at a time (the generator's temporaries are O(n) float64)
    x = torch.empty((C, n), dtype=torch.float32, device=dev)
    piece = 10 * 60 * SR
    for p0 in range(0, n, piece):
        p1 = min(n, p0 + piece)
        x[:, p0:p1] = synth_device(torch, C, p1 - p0, rank * C + p0 // piece, dev)
    transport = os.environ.get("B200GATE_GATHER", "store") if world > 1 else "none"
    reserve = int(os.environ.get("B200GATE_RESERVE_SMS", {2: "12", 4: "16"}.get(world, "32")))
    use_peer = transport == "peer"
    dg = DeviceGate(sr=SR, stationary=True, n_fft=1024, hop_length=256, chunk_size=cs, padding=30000,
                    reserve_sms=(reserve if transport == "store" else 16 if transport == "nccl" else 0), workspace_limit_bytes=48e9)
    if world == 1:
        dg.noise_stats(x)
    else:
        chained_noise_stats(dg, x, rank, world)
    comm = torch.cuda.Stream()
    pg = ps = None
    if use_peer:
        from noisereduce_b200.parallel import PeerGather
        pg = PeerGather(world, rank, (2, world, C, args.slab_chunks * cs), torch.float32, dev)
    elif transport == "store":
        from noisereduce_b200.parallel import PeerStore
        ps = PeerStore(world, rank, (2, world, C, args.slab_chunks * cs), torch.float32, dev)
    buffers = make_slab_ring(C, args.slab_chunks * cs, world, torch.float32, dev) if (pg is None and ps is None) else None
    sums = []

    def consume(g, first, si):
        return g.sum(dtype=torch.float64)            # stays on the device; read after the timed region

    def step():
        if ps is not None:
            from noisereduce_b200.parallel import slab_ring_peer_store
            return slab_ring_peer_store(dg, x, world, args.slab_chunks, consume, ps, push_ctas=reserve)
        return sharded_run_slab_ring(dg, x, world, comm, slab_chunks=args.slab_chunks, consume=consume, buffers=buffers, peer=pg)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        sums = step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1) / args.steps
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    checks = torch.stack(sums).cpu()
    agree = True
    if world > 1:                                     # every rank consumed the same gathered slabs
        mine = checks.to(dev)
        allc = torch.empty((world, mine.numel()), dtype=mine.dtype, device=dev)
        dist.all_gather_into_tensor(allc.view(-1), mine)
        agree = bool((allc == allc[0:1]).all().item())
    if rank == 0:
        gathered_bytes = (world - 1) * C * n * 4
        line = {"metric": METRIC, "value": world * C * n / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": f"{world * C}ch x {args.minutes:g}min synthetic 48kHz stationary, chunk_size=60s, "
                                       f"channel-sharded 64 ch/GPU, slab-ring all-gather (configs[4])",
                           "channels_per_gpu": C, "samples_per_channel": n, "chunk_size": cs, "padding": 30000,
                           "slab_chunks": args.slab_chunks, "shard_bytes_in_hbm": C * n * 4,
                           "gather_transport": {"store": "kernel-nvlink-stores", "peer": "peer-copy-engine", "nccl": "nccl", "none": "none"}[transport]},
                "all_gather_bytes_received_per_rank": gathered_bytes,
                "all_gather_GBps_per_rank": gathered_bytes / (ms * 1e-3) / 1e9,
                "slab_checksum_total": float(checks.sum()), "slabs": len(sums), "checksums_agree_across_ranks": agree,
                "peak_memory_GB": torch.cuda.max_memory_allocated() / 1e9, "stats_last_slab": dg.gate.stats()}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
