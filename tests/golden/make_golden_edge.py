"""Generate tests/golden/edge_cases.npz by running the UNMODIFIED reference on tests/golden/edge_cases.py's inputs.
Build container only (needs /root/reference):   python tests/golden/make_golden_edge.py"""
import os
import sys
import warnings

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(HERE, "..", ".."))
warnings.filterwarnings("ignore")

import scipy  # noqa: E402
import noisereduce as nr  # noqa: E402  (the reference)
from tests.golden.edge_cases import SR, cases, refused  # noqa: E402

out = {"versions": np.array([f"reference=51c8534 v3.0.3 numpy={np.__version__} scipy={scipy.__version__}"])}
for name, (y, kw) in cases().items():
    out[name] = nr.reduce_noise(y=y, sr=SR, **kw)
    print(f"{name:40s} {out[name].dtype} {out[name].shape}")
for name, (y, kw, exc, msg) in refused().items():
    try:
        nr.reduce_noise(y=y, sr=SR, **kw)
    except exc as e:
        print(f"{name:40s} raises {type(e).__name__}: {str(e)[:70]}")
    else:
        raise SystemExit(f"the reference accepted {name}")
np.savez_compressed(os.path.join(HERE, "edge_cases.npz"), **out)
