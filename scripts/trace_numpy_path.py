"""Where reduce_noise(y=pageable float32[64, 28.8M]) spends its time on the GPU box: handle creation, noise statistics,
the run (staging + H2D + kernels + D2H into the pooled result), teardown.   python scripts/trace_numpy_path.py [reps]"""
import gc
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from noisereduce_b200 import _cabi  # noqa: E402
from noisereduce_b200.spectralgate.stationary import SpectralGateStationary  # noqa: E402
import noisereduce_b200 as nrb  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 4
SR, C, n = 48000, 64, 28_800_000
rng = np.random.default_rng(0)
y = np.empty((C, n), np.float32)
for c in range(C):
    y[c] = 0.05 * rng.standard_normal(n, dtype=np.float32)
kw = dict(y_noise=None, prop_decrease=1.0, n_std_thresh_stationary=1.5, chunk_size=600000, clip_noise_stationary=True,
          padding=30000, n_fft=1024, win_length=None, hop_length=256, time_constant_s=2.0, freq_mask_smooth_hz=500,
          time_mask_smooth_ms=50, tmp_folder=None, use_tqdm=False, n_jobs=1)
nrb.reduce_noise(y=y[:2], sr=SR, stationary=True, n_fft=1024, hop_length=256)
for rep in range(reps):
    t0 = time.perf_counter()
    sg = SpectralGateStationary(y=y, sr=SR, **kw)
    t1 = time.perf_counter()
    res = sg.get_traces()
    t2 = time.perf_counter()
    st = sg._gate.stats()
    del sg
    gc.collect()
    t3 = time.perf_counter()
    del res
    gc.collect()
    t4 = time.perf_counter()
    print(json.dumps({"rep": rep, "construct+noise_stats_ms": round((t1 - t0) * 1e3, 1), "get_traces_ms": round((t2 - t1) * 1e3, 1),
                      "destroy_handle_ms": round((t3 - t2) * 1e3, 1), "drop_result_ms": round((t4 - t3) * 1e3, 1),
                      "lib_last_run_ms": round(st["last_run_ms"], 1), "lib_h2d_ms": round(st["last_h2d_ms"], 1),
                      "lib_d2h_ms": round(st["last_d2h_ms"], 1), "leased_bytes": _cabi._pinned_leased_bytes}), flush=True)
