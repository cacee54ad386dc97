"""Round-1 experiment (GPU): (1) does a batch small enough to keep the cached spectra L2-resident beat one big
batch?  (2) PCIe bandwidths and the e2e slab size.  Prints one line per setting."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import synth_device, SR, C_PER_GPU  # noqa: E402
from noisereduce_b200.device import DeviceGate  # noqa: E402

dev = torch.device("cuda", 0)
n = 10 * 60 * SR
C = C_PER_GPU
x = synth_device(torch, C, n, 0, dev)
out = torch.empty_like(x)


def timed(dg, steps=3):
    dg.run(x, out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        dg.run(x, out)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps, dg.gate.stats()


if "l2" in sys.argv[1:] or len(sys.argv) == 1:
    for ws in (64e9, 2e9, 500e6, 240e6, 120e6, 80e6, 56e6, 28e6):
        dg = DeviceGate(sr=SR, stationary=True, n_fft=1024, hop_length=256, workspace_limit_bytes=ws)
        dg.noise_stats(x)
        ms, s = timed(dg)
        print(f"ws={ws:.3g} step={ms:.2f} ms k1={s['k1_ms']:.2f} sm={s['smooth_ms']:.2f} k2={s['k2_ms']:.2f} "
              f"launches={s['kernel_launches']}", flush=True)
        del dg

if "pcie" in sys.argv[1:] or len(sys.argv) == 1:
    nb = 1 << 30
    h1 = torch.empty(nb, dtype=torch.uint8, pin_memory=True)
    h2 = torch.empty(nb, dtype=torch.uint8, pin_memory=True)
    d1 = torch.empty(nb, dtype=torch.uint8, device=dev)
    d2 = torch.empty(nb, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def bw(fn, reps=4):
        fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        return reps * nb / (time.perf_counter() - t0) / 1e9

    def h2d():
        with torch.cuda.stream(s1):
            d1.copy_(h1, non_blocking=True)

    def d2h():
        with torch.cuda.stream(s2):
            h2.copy_(d2, non_blocking=True)

    def both():
        h2d()
        d2h()

    print(f"pcie h2d {bw(h2d):.1f} GB/s  d2h {bw(d2h):.1f} GB/s  duplex {bw(both):.1f} GB/s each", flush=True)
    del h1, h2, d1, d2

if "slab" in sys.argv[1:] or len(sys.argv) == 1:
    hx = torch.empty((C, n), dtype=torch.float32, pin_memory=True)
    hy = torch.empty((C, n), dtype=torch.float32, pin_memory=True)
    hx.copy_(x)
    for mb in (160, 320, 512, 1024, 2048):
        os.environ["B200GATE_SLAB_MB"] = str(mb)
        dg = DeviceGate(sr=SR, stationary=True, n_fft=1024, hop_length=256, workspace_limit_bytes=64e9)
        dg.noise_stats(x)
        g = dg.gate

        def e2e():
            g._check(g.lib.dll.b200gate_run(g._h, hx.data_ptr(), hy.data_ptr(), 0, C, n, n, n, 0, None))

        e2e()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            e2e()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 3
        print(f"slab={mb} MB e2e {dt*1e3:.1f} ms = {C*n/dt/1e9:.2f} Gsamples/s", flush=True)
        del dg
