"""Throughput of the general-geometry (float64) family on a few off-default STFT geometries, device-resident."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import synth_device  # noqa: E402
from noisereduce_b200.device import DeviceGate  # noqa: E402

dev = torch.device("cuda", 0)
C, sr = 16, 16000
n = 10 * 60 * sr
x = synth_device(torch, C, n, 0, dev)
out = torch.empty_like(x)
for stationary in (True, False):
    for geo in (dict(n_fft=512), dict(n_fft=512, win_length=400, hop_length=100), dict(n_fft=2048),
                dict(n_fft=1024, path_flags=4), dict(n_fft=1024)):
        if geo.get("n_fft") == 2048 and not stationary:
            continue
        dg = DeviceGate(sr=sr, stationary=stationary, **geo)
        if stationary:
            dg.noise_stats(x)
        dg.run(x, out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            dg.run(x, out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        print(f"stationary={stationary} {geo}: {ms:.2f} ms  {C * n / ms / 1e6:.2f} Gsamples/s", flush=True)
        del dg
