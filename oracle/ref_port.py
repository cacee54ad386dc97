"""Performance-faithful CPU port of the reference path (TEST / BASELINE INFRASTRUCTURE).

oracle/spectral_gate_oracle.py restates the arithmetic explicitly; this module instead keeps the
reference's *cost structure* so that bench.py's ``cpu_baseline`` / ``--impl reference`` legs time
what a user of timsainb/noisereduce actually runs on the host: the same scipy.signal calls
(stft / fftconvolve / istft / filtfilt, as at noisereduce/spectralgate/stationary.py:87-125 and
nonstationary.py:51-95), the same float64 promotion of every padded chunk (base.py:130-142), a Python
loop over channels inside a chunk (stationary.py:86) and joblib process-parallelism over chunks
(base.py:206-216).  The reference itself (/root/reference) is not present on the GPU box, so it
cannot be timed there directly; tests/test_oracle_golden.py pins this port to the reference's outputs.
Only bench.py and tests/ may import this module.
"""
import numpy as np
from joblib import Parallel, delayed
from scipy.signal import fftconvolve, filtfilt, istft, stft

from . import spectral_gate_oracle as sgo


def _unit_stationary(x, thresh, p, filt, N, W, H):
    _, _, X = stft(x, nfft=N, noverlap=W - H, nperseg=W, padded=False)
    db = sgo.amp_to_db(X)
    mask = (db > thresh[:, None]) * p + np.ones(db.shape) * (1.0 - p)
    if filt is not None:
        mask = fftconvolve(mask, filt, mode="same")
    _, y = istft(X * mask, nfft=N, noverlap=W - H, nperseg=W)
    return y


def _unit_nonstationary(x, b, n_mult, slope, p, filt, N, W, H):
    _, _, X = stft(x, nfft=N, noverlap=W - H, nperseg=W, padded=False)
    A = np.abs(X)
    S = filtfilt([b], [1, b - 1], A, axis=-1, padtype=None)
    with np.errstate(divide="ignore", invalid="ignore"):
        mask = sgo.sigmoid((A - S) / S, -n_mult, slope)
    if filt is not None:
        mask = fftconvolve(mask, filt, mode="same")
    mask = mask * p + np.ones(mask.shape) * (1.0 - p)
    _, y = istft(X * mask, nfft=N, noverlap=W - H, nperseg=W)
    return y


def _chunk_job(y2d, i1, i2, lo, hi, stationary, args):
    chunk = sgo.read_chunk(y2d, i1, i2)
    out = np.zeros(chunk.shape)
    for c in range(chunk.shape[0]):
        yc = _unit_stationary(chunk[c], *args) if stationary else _unit_nonstationary(chunk[c], *args)
        out[c, : len(yc)] = yc
    return out[:, lo - i1: hi - i1]


def reduce_noise(y, sr, cfg: sgo.GateConfig, y_noise=None, n_jobs=1):
    """Same result as oracle.spectral_gate_oracle.reduce_noise / the reference's reduce_noise."""
    y2d, flat = sgo._as_2d(y)
    C, n = y2d.shape
    N, W, H = cfg.resolve()
    smooth, nf, nt = sgo.smoothing_extents(sr, N, H, cfg.freq_mask_smooth_hz, cfg.time_mask_smooth_ms)
    filt = sgo.smoothing_filter(nf, nt) if smooth else None
    if cfg.stationary:
        yn2d = y2d if y_noise is None else sgo._as_2d(y_noise)[0]
        yn = sgo.collapse_noise(yn2d, cfg.chunk_size, cfg.clip_noise_stationary)
        _, _, Xn = stft(yn, nfft=N, noverlap=W - H, nperseg=W, padded=False)
        db = sgo.amp_to_db(Xn)
        thresh = np.mean(db, axis=1) + np.std(db, axis=1) * cfg.n_std_thresh_stationary
        args = (thresh, cfg.prop_decrease, filt, N, W, H)
    else:
        b = sgo.iir_coefficient(cfg.time_constant_s, sr, H)
        args = (b, cfg.thresh_n_mult_nonstationary, cfg.sigmoid_slope_nonstationary, cfg.prop_decrease, filt, N, W, H)
    table = sgo.chunk_table(n, cfg.chunk_size, cfg.padding)
    parts = Parallel(n_jobs=n_jobs)(
        delayed(_chunk_job)(y2d, i1, i2, lo, hi, cfg.stationary, args) for (i1, i2, lo, hi) in table
    )
    out = np.zeros((C, n))
    for (i1, i2, lo, hi), part in zip(table, parts):
        out[:, lo:hi] = part
    res = sgo.cast_like_reference(out, y2d.dtype)
    return res.flatten() if flat else res
