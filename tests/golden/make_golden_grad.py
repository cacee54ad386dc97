"""Generate tests/golden/torchgate_grad.npz from the UNMODIFIED reference (build container only): forward output and
the gradient of sum(y * g) with respect to x through noisereduce.torchgate.TorchGate on CPU (torchgate.py:200-264: masks
under no_grad, gradient through stft -> * mask -> istft).   python tests/golden/make_golden_grad.py"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
from noisereduce.torchgate import TorchGate  # noqa: E402
from tests.synth_host import synth_torchgate  # noqa: E402

sr = 16000
x = synth_torchgate(B=2, n=24 * 256 + 100)                 # not a multiple of hop: the last 100 samples lie beyond the output
gen = torch.Generator().manual_seed(11)
out = {"x": x, "sr": sr, "versions": f"torch {torch.__version__} numpy {np.__version__}"}
for name, kw in (("stat", {}), ("nonstat", dict(nonstationary=True, prop_decrease=0.8))):
    tg = TorchGate(sr=sr, **kw)
    xt = torch.from_numpy(x).clone().requires_grad_(True)
    y = tg(xt)
    g = torch.randn(y.shape, generator=gen)
    (y * g).sum().backward()
    out[f"{name}_y"] = y.detach().numpy()
    out[f"{name}_g"] = g.numpy()
    out[f"{name}_grad"] = xt.grad.numpy()
np.savez_compressed(os.path.join(HERE, "torchgate_grad.npz"), **out)
print({k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items()})
