"""Build-container tool (needs /root/reference and the CPU simulator build; never runs on the GPU box): random
reduce_noise() configurations -- sample rate, n_fft / win / hop, chunking, smoothing extents, dtypes, shapes -- through
noisereduce_b200 on the simulator library against the UNMODIFIED reference.

    python scripts/fuzz_vs_reference.py [seed] [cases]

Reports shape / dtype / exception-type mismatches, integer outputs off by more than 1 LSB, NaN patterns that differ and
float outputs beyond 1e-4 rel-inf.  (A stationary case may exceed 1e-4 locally when one mask decision flips on the
threshold caveat of DESIGN.md section 5: with the reference's thresholds injected the outputs are equal.)
Round 2: seeds 1-3, 160 cases: 0 mismatches apart from one such flip (2.2e-4 over 1329 samples, 0.0 with injected thresholds)."""
import sys, warnings, time
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__)))); sys.path.insert(0, "/root/reference")
import numpy as np
warnings.filterwarnings("ignore")
import noisereduce as ref_nr
from noisereduce_b200 import _cabi
from tests.cusim_util import cusim_library
_cabi._LIB = cusim_library()
import noisereduce_b200 as nr
from tests.synth_host import synth_small
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
base = synth_small(C=3, n=60000)
bad = 0
N = int(sys.argv[2]) if len(sys.argv) > 2 else 30
for it in range(N):
    sr = int(rng.choice([8000, 16000, 22050, 44100, 48000]))
    n_fft = int(rng.choice([256, 512, 1024, 1024, 1024, 2048]))
    C = int(rng.integers(1, 4)); n = int(rng.integers(n_fft + 10, 30000))
    stat = bool(rng.integers(0, 2))
    kw = dict(sr=sr, n_fft=n_fft, stationary=stat)
    if rng.random() < 0.5:
        kw["chunk_size"] = int(rng.integers(max(n_fft, 600), 12000)); kw["padding"] = int(rng.integers(0, 3000))
    if rng.random() < 0.3: kw["prop_decrease"] = float(rng.choice([0.0, 0.3, 0.8, 1.0]))
    if rng.random() < 0.3: kw["hop_length"] = int(n_fft // rng.choice([2, 4, 8]))
    if rng.random() < 0.2: kw["win_length"] = int(n_fft * rng.choice([0.5, 0.75, 1.0]))
    if rng.random() < 0.3: kw["freq_mask_smooth_hz"] = float(rng.choice([100, 300, 500, 1000, 2000]))
    if rng.random() < 0.3: kw["time_mask_smooth_ms"] = float(rng.choice([20, 50, 100, 200]))
    if stat and rng.random() < 0.3: kw["n_std_thresh_stationary"] = float(rng.choice([0.5, 1.5, 3.0]))
    if not stat and rng.random() < 0.3: kw["time_constant_s"] = float(rng.choice([0.1, 0.5, 2.0, 4.0]))
    dt = rng.choice(["f32", "f32", "f64", "i16"])
    y = base[:C, :n] if C > 1 or rng.random() < 0.5 else base[0, :n]
    if dt == "f64": y = y.astype(np.float64)
    if dt == "i16": y = (y * 20000).astype(np.int16)
    if "win_length" in kw and "hop_length" in kw and kw["hop_length"] > kw["win_length"]: kw.pop("hop_length")
    try: r = ref_nr.reduce_noise(y=y, **kw); re_ = None
    except Exception as e: r, re_ = None, e
    try: o = nr.reduce_noise(y=y, **kw); oe = None
    except Exception as e: o, oe = None, e
    tag = f"{it:3d} {dt} C={C} n={n} " + " ".join(f"{k}={v}" for k, v in kw.items())
    if re_ or oe:
        same = re_ is not None and oe is not None and type(re_).__name__ == type(oe).__name__
        if not same:
            bad += 1; print("MISMATCH-EXC", tag, "| ref:", repr(re_)[:100], "| ours:", repr(oe)[:100])
        continue
    if o.shape != r.shape or o.dtype != r.dtype:
        bad += 1; print("MISMATCH-SHAPE", tag, o.shape, r.shape, o.dtype, r.dtype); continue
    if np.issubdtype(r.dtype, np.integer):
        e = int(np.abs(o.astype(np.int64) - r.astype(np.int64)).max())
        if e > 1: bad += 1; print("MISMATCH-LSB", e, tag)
    else:
        if not np.array_equal(np.isnan(o), np.isnan(r)): bad += 1; print("MISMATCH-NAN", tag); continue
        m = float(np.abs(np.nan_to_num(r)).max()) or 1.0
        e = float(np.abs(np.nan_to_num(o) - np.nan_to_num(r)).max()) / m
        if e > 1e-4: bad += 1; print(f"MISMATCH-VAL {e:.2e}", tag)
print("cases", N, "mismatches", bad)
