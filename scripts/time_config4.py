"""Config 4 timing (informational): TorchGate batch=256 x 10 s @ 16 kHz on one B200, CUDA events."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from noisereduce_b200.torchgate import TorchGate
g = torch.Generator(device="cuda").manual_seed(1234)
x = 0.05 * torch.randn((256, 160000), device="cuda", generator=g)
res = {}
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
    res[name] = {"ms_per_forward": ms, "samples_per_s": 256 * 160000 / (ms * 1e-3)}
print(json.dumps({"config": "TorchGate batch=256 x 10 s @ 16 kHz, 1 x B200", **res}))
