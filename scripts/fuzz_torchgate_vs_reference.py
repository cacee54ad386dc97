"""Build-container tool (needs /root/reference and the CPU simulator build): random TorchGate configurations (sample rate, n_fft /
hop, smoothing extents, stationary / moving-mean gate, float32 / float64, batch and length, optional xn) through
noisereduce_b200.TorchGate on the simulator library against the UNMODIFIED reference module on CPU.
    python scripts/fuzz_torchgate_vs_reference.py
Round 2: 30 cases, 0 mismatches (shape, dtype, exception, 1e-4 rel-inf)."""
import sys, warnings
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__)))); sys.path.insert(0, "/root/reference")
import numpy as np, torch
warnings.filterwarnings("ignore")
from noisereduce.torchgate import TorchGate as RefTG
from noisereduce_b200.torchgate import TorchGate as OurTG
from tests.cusim_util import cusim_library
from tests.synth_host import synth_torchgate
lib = cusim_library()
rng = np.random.default_rng(21)
x = torch.from_numpy(synth_torchgate(B=4, n=30000))
bad = 0
for it in range(30):
    sr = int(rng.choice([8000, 16000, 22050, 44100, 48000])); n_fft = int(rng.choice([256, 512, 1024, 1024, 2048]))
    kw = dict(sr=sr, n_fft=n_fft, nonstationary=bool(rng.integers(0, 2)))
    if rng.random() < 0.3: kw["prop_decrease"] = float(rng.choice([0.0, 0.4, 1.0]))
    if rng.random() < 0.3: kw["hop_length"] = int(n_fft // rng.choice([2, 4, 8]))
    if rng.random() < 0.3: kw["freq_mask_smooth_hz"] = float(rng.choice([200, 500, 1500]))
    if rng.random() < 0.3: kw["time_mask_smooth_ms"] = float(rng.choice([30, 50, 120]))
    if rng.random() < 0.3: kw["n_std_thresh_stationary"] = float(rng.choice([0.5, 1.5, 2.5]))
    if rng.random() < 0.3: kw["n_movemean_nonstationary"] = int(rng.choice([5, 20, 41]))
    B = int(rng.integers(1, 5)); n = int(rng.integers(2 * n_fft, 30000))
    xx = x[:B, :n]
    if rng.random() < 0.25: xx = xx.double()
    xn = None
    if not kw["nonstationary"] and rng.random() < 0.3: xn = x[:B, 100:100 + int(rng.integers(2 * n_fft, 20000))].to(xx.dtype)
    try: r = RefTG(**kw)(xx, xn); re_ = None
    except Exception as e: r, re_ = None, e
    try: o = OurTG(**kw)(xx, xn, _lib=lib); oe = None
    except Exception as e: o, oe = None, e
    tag = f"{it} B={B} n={n} {xx.dtype} xn={None if xn is None else tuple(xn.shape)} " + " ".join(f"{k}={v}" for k, v in kw.items())
    if re_ or oe:
        if not (re_ is not None and oe is not None): bad += 1; print("EXC-MISMATCH", tag, "|", repr(re_)[:90], "|", repr(oe)[:90])
        continue
    if o.shape != r.shape or o.dtype != r.dtype: bad += 1; print("SHAPE", tag, o.shape, r.shape, o.dtype, r.dtype); continue
    m = float(r.abs().max()) or 1.0; e = float((o - r).abs().max()) / m
    if e > 1e-4: bad += 1; print(f"VAL {e:.2e}", tag)
print("mismatches", bad)
