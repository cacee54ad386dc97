"""reduce_noise() on a pageable float32[64, 28.8 M] array against the number of host staging workers (B200GATE_HOST_THREADS is read
on every run): python scripts/sweep_host_threads.py [counts ...]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import noisereduce_b200 as nrb  # noqa: E402

counts = [int(a) for a in sys.argv[1:]] or [8, 12, 16, 24, 32, 48]
SR, C, n = 48000, 64, 28_800_000
rng = np.random.default_rng(0)
row = (0.05 * rng.standard_normal(n, dtype=np.float32))
y = np.empty((C, n), np.float32)
for c in range(C):
    y[c] = np.roll(row, 1000 * c)
nrb.reduce_noise(y=y[:2], sr=SR, stationary=True, n_fft=1024, hop_length=256)
res = nrb.reduce_noise(y=y, sr=SR, stationary=True, n_fft=1024, hop_length=256)      # fills the pools
del res
for rep in range(2):
    for k in counts:
        os.environ["B200GATE_HOST_THREADS"] = str(k)
        ts = []
        for _ in range(3):
            t0 = time.perf_counter()
            res = nrb.reduce_noise(y=y, sr=SR, stationary=True, n_fft=1024, hop_length=256)
            ts.append((time.perf_counter() - t0) * 1e3)
            del res
        print(json.dumps({"host_threads": k, "rep": rep, "ms_per_call": [round(t, 1) for t in ts]}), flush=True)
