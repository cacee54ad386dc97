"""Seeded host-side synthetic signals shared by the golden generator and the tests."""
import numpy as np


def synth_small(C=2, n=30000, sr=16000, seed=7):
    """White noise floor + gated tones (a different pitch per channel) + one click: gives a
    non-trivial mask, exercises chunk seams, and has loud and quiet stretches."""
    rng = np.random.default_rng(seed)
    t = np.arange(n) / sr
    y = 0.05 * rng.standard_normal((C, n))
    for c in range(C):
        f0 = 440.0 * 2 ** (c / 3)
        gate = ((t % 0.5) < 0.2).astype(np.float64)
        y[c] += 0.25 * gate * np.sin(2 * np.pi * f0 * t) + 0.1 * gate * np.sin(2 * np.pi * 3.1 * f0 * t)
    y[0, n // 3] += 0.9
    return y.astype(np.float32)


def synth_torchgate(B=3, n=24000, sr=16000, seed=11):
    rng = np.random.default_rng(seed)
    t = np.arange(n) / sr
    x = 0.05 * rng.standard_normal((B, n))
    for b in range(B):
        x[b] += 0.2 * np.sin(2 * np.pi * (500 + 300 * b) * t) * ((t % 0.6) < 0.25)
    return x.astype(np.float32)


def band_noise(n, sr, lo=2000.0, hi=12000.0, seed=3):
    """Seeded band-limited noise for the scenarios of the reference's own test file (test_reduction.py builds its input
    as `fish + band_limited_noise(2000, 12000) * 10`): random phases on the rfft bins inside [lo, hi], unit peak."""
    rng = np.random.default_rng(seed)
    f = np.fft.rfftfreq(n, 1.0 / sr)
    spec = np.where((f >= lo) & (f <= hi), np.exp(2j * np.pi * rng.random(f.shape[0])), 0.0)
    x = np.fft.irfft(spec, n)
    return x / np.abs(x).max()
