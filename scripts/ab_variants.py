"""Config-2 step / per-kernel times of library variants built by scripts/build_variant.py.
    python scripts/ab_variants.py default v12 ...      (names; 'default' = libb200gate.so)   [--flags N]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import synth_device, SR, C_PER_GPU  # noqa: E402
from noisereduce_b200 import _cabi  # noqa: E402
from noisereduce_b200.device import DeviceGate  # noqa: E402

names = [a for a in sys.argv[1:] if not a.startswith("--")] or ["default"]
flags = int(sys.argv[sys.argv.index("--flags") + 1]) if "--flags" in sys.argv else 0
dev = torch.device("cuda", 0)
n = 10 * 60 * SR
x = synth_device(torch, C_PER_GPU, n, 0, dev)
base = None
for rep in range(2):
    for nm in names:
        path = _cabi.DEFAULT_LIB if nm == "default" else os.path.join(ROOT, "noisereduce_b200", f"libb200gate_{nm}.so")
        lib = _cabi.GateLibrary(path)
        dg = DeviceGate(sr=SR, stationary=True, n_fft=1024, hop_length=256, workspace_limit_bytes=64e9, path_flags=flags, lib=lib)
        dg.noise_stats(x)
        out = torch.empty_like(x)
        for _ in range(3):
            dg.run(x, out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            dg.run(x, out)
        e1.record()
        torch.cuda.synchronize()
        s = dg.gate.stats()
        eq = ""
        if base is None:
            base = out[:4].clone()
        else:
            eq = f"  max|diff vs first| {float((out[:4] - base).abs().max()):.3e}"
        print(f"rep {rep} {nm}: step {e0.elapsed_time(e1) / 5:.2f} ms  k1 {s['k1_ms']:.2f}  smooth {s['smooth_ms']:.2f}  k2 {s['k2_ms']:.2f}{eq}", flush=True)
        del dg
