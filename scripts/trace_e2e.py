"""Slab timeline of the host-buffer path at config 2 (B200GATE_TRACE=1 prints per-slab event times)."""
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
hx = torch.empty((C, n), dtype=torch.float32, pin_memory=True)
hy = torch.empty((C, n), dtype=torch.float32, pin_memory=True)
hx.copy_(x)
dg = DeviceGate(sr=SR, stationary=True, n_fft=1024, hop_length=256, workspace_limit_bytes=64e9)
dg.noise_stats(x)
g = dg.gate
for it in range(3):
    if it == 2:
        os.environ["B200GATE_TRACE"] = "1"
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    g._check(g.lib.dll.b200gate_run(g._h, hx.data_ptr(), hy.data_ptr(), 0, C, n, n, n, 0, None))
    torch.cuda.synchronize()
    print(f"iter {it}: {(time.perf_counter() - t0) * 1e3:.1f} ms", flush=True)
