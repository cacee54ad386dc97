"""Device-resident timing of BASELINE configs 2, 3 (64 ch x 10 min @ 48 kHz) and 4 (TorchGate 256 x 10 s @ 16 kHz) on one
B200: CUDA events, per-kernel stage times from the library.   python scripts/time_configs.py [2] [3] [4]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import synth_device, SR, C_PER_GPU  # noqa: E402
from noisereduce_b200.device import DeviceGate  # noqa: E402

which = [int(a) for a in sys.argv[1:]] or [2, 3, 4]
dev = torch.device("cuda", 0)
res = {}
if 2 in which or 3 in which:
    n = 10 * 60 * SR
    x = synth_device(torch, C_PER_GPU, n, 0, dev)
    out = torch.empty_like(x)
    for cfg in (2, 3):
        if cfg not in which:
            continue
        kw = dict(stationary=True, n_fft=1024, hop_length=256) if cfg == 2 else dict(stationary=False, n_fft=2048)
        dg = DeviceGate(sr=SR, workspace_limit_bytes=64e9, **kw)
        if cfg == 2:
            dg.noise_stats(x)
        for _ in range(2):
            dg.run(x, out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        K = 5
        for _ in range(K):
            dg.run(x, out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / K
        s = dg.gate.stats()
        res[f"config{cfg}"] = {"ms_per_step": ms, "samples_per_s": C_PER_GPU * n / (ms * 1e-3),
                               "k1_ms": s["k1_ms"], "mid_ms": s["smooth_ms"], "k2_ms": s["k2_ms"],
                               "launches": s["kernel_launches"], "finite": bool(torch.isfinite(out[:2]).all().item())}
        print(json.dumps({f"config{cfg}": res[f"config{cfg}"]}), flush=True)
        del dg
    del x, out
if 4 in which:
    from noisereduce_b200.torchgate import TorchGate
    g = torch.Generator(device="cuda").manual_seed(1234)
    x = 0.05 * torch.randn((256, 160000), device="cuda", generator=g)
    for name, kw in (("stationary", {}), ("nonstationary", {"nonstationary": True})):
        tg = TorchGate(sr=16000, **kw).to("cuda")
        for _ in range(3):
            y = tg(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            y = tg(x)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        res[f"config4_{name}"] = {"ms_per_forward": ms, "samples_per_s": 256 * 160000 / (ms * 1e-3)}
        print(json.dumps({f"config4_{name}": res[f"config4_{name}"]}), flush=True)
print(json.dumps(res))
