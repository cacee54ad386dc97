"""CPU oracle for the TorchGate surface (TEST INFRASTRUCTURE -- see oracle/__init__.py).

Restates noisereduce/torchgate/torchgate.py:200-264 (forward), :127-165 (_stationary_mask),
:168-198 (_nonstationary_mask), :74-124 (smoothing filter) and torchgate/utils.py:6-66 in numpy.
torch.stft / torch.istft / conv1d / conv2d / std_mean are third-party (torch 2.11.0 in the build
container); their documented algorithms are written out here.

Arithmetic is float64 throughout, except for the tables the reference itself builds in float32
(the Hann window from torch.hann_window and the smoothing taps from torch.linspace); those can be
injected so the pinning tests can use the very tables the reference used.
"""
from __future__ import annotations

import math
from typing import Optional

import numpy as np

from . import spectral_gate_oracle as sgo

EPS64 = sgo.EPS64


def hann_window_f32(W: int) -> np.ndarray:
    """torch.hann_window(W) (periodic, float32): cos of a float32-rounded angle, then -0.5*c + 0.5.
    Matches torch to 1 ulp(float32); pass the real torch table for bit-level pinning."""
    a = np.arange(W, dtype=np.float32) * np.float32(math.pi * 2.0 / W)
    return (np.cos(a) * np.float32(-0.5) + np.float32(0.5)).astype(np.float32)


def smoothing_filter_f32(n_grad_freq: int, n_grad_time: int) -> np.ndarray:
    """torchgate.py:107-124: float32 triangles from torch.linspace, outer product, / sum."""
    vf = sgo.triangle_taps(n_grad_freq).astype(np.float32)
    vt = sgo.triangle_taps(n_grad_time).astype(np.float32)
    filt = np.outer(vf, vt).astype(np.float32)
    return (filt / filt.sum(dtype=np.float32)).astype(np.float32)


def smoothing_extents(sr, n_fft, hop_length, freq_mask_smooth_hz, time_mask_smooth_ms):
    """torchgate.py:85-105.  Same integer arithmetic as the numpy surface."""
    if freq_mask_smooth_hz is None and time_mask_smooth_ms is None:
        return False, 1, 1
    nf = 1 if freq_mask_smooth_hz is None else int(freq_mask_smooth_hz / (sr / (n_fft / 2)))
    if nf < 1:
        raise ValueError("freq_mask_smooth_hz needs to be at least {} Hz".format(int(sr / (n_fft / 2))))
    nt = 1 if time_mask_smooth_ms is None else int(time_mask_smooth_ms / ((hop_length / sr) * 1000))
    if nt < 1:
        raise ValueError("time_mask_smooth_ms needs to be at least {} ms".format(int((hop_length / sr) * 1000)))
    if nf == 1 and nt == 1:
        return False, 1, 1
    return True, nf, nt


def stft_center(x: np.ndarray, n_fft: int, win_length: int, hop_length: int, window: np.ndarray) -> np.ndarray:
    """torch.stft(x, n_fft, hop, win_length, window, center=True, pad_mode='constant',
    normalized=False, onesided=True, return_complex=True) (torchgate.py:223-232).
    x: [B, L] -> X: [B, F, T], T = 1 + L // hop.  Window centre-padded to n_fft."""
    B, L = x.shape
    N, W, H = n_fft, win_length, hop_length
    wfull = np.zeros(N)
    left = (N - W) // 2
    wfull[left: left + W] = window.astype(np.float64)
    xe = np.concatenate([np.zeros((B, N // 2)), x.astype(np.float64), np.zeros((B, N // 2))], axis=1)
    T = 1 + L // H
    idx = (np.arange(T) * H)[:, None] + np.arange(N)[None, :]
    frames = xe[:, idx] * wfull[None, None, :]                      # [B, T, N]
    X = sgo.rfft(frames, N)                                         # [B, T, F]
    return np.ascontiguousarray(np.transpose(X, (0, 2, 1)))


def istft_center(Y: np.ndarray, n_fft: int, win_length: int, hop_length: int, window: np.ndarray) -> np.ndarray:
    """torch.istft(Y, n_fft, hop, win_length, window, center=True) (torchgate.py:255-262):
    irfft * window, overlap-add, divide by overlap-added window^2, trim n_fft//2 both sides."""
    B, F, T = Y.shape
    N, W, H = n_fft, win_length, hop_length
    wfull = np.zeros(N)
    left = (N - W) // 2
    wfull[left: left + W] = window.astype(np.float64)
    seg = sgo.irfft(np.ascontiguousarray(np.transpose(Y, (0, 2, 1))), N) * wfull   # [B, T, N]
    L = N + (T - 1) * H
    out = np.zeros((B, L))
    env = np.zeros(L)
    for t in range(T):
        out[:, t * H: t * H + N] += seg[:, t]
        env[t * H: t * H + N] += wfull * wfull
    out = out[:, N // 2: L - N // 2]
    env = env[N // 2: L - N // 2]
    return out / env[None, :]


def amp_to_db(X: np.ndarray, top_db: float = 40.0, eps: float = EPS64) -> np.ndarray:
    """torchgate/utils.py:6-23 (top_db = 40; floor = max over time per (batch, freq) - top_db)."""
    x_db = 20 * np.log10(np.abs(X) + eps)
    return np.maximum(x_db, x_db.max(axis=-1, keepdims=True) - top_db)


def conv2d_same(mask: np.ndarray, filt: np.ndarray) -> np.ndarray:
    """conv2d(mask[:,None], filt[None,None], padding='same') (torchgate.py:244-249): zero-padded
    cross-correlation; the kernel is symmetric so it equals convolution.  mask: [B, F, T]."""
    return np.stack([sgo.conv2d_same_zero(m, filt.astype(np.float64)) for m in mask])


def moving_mean_same(A: np.ndarray, n: int) -> np.ndarray:
    """conv1d(A, ones(n), padding='same') / n (torchgate.py:179-190).  PyTorch 'same' padding for an
    even kernel pads (n-1)//2 on the left and n-1-(n-1)//2 on the right (left 9 / right 10 for 20)."""
    left = (n - 1) // 2
    right = n - 1 - left
    pad = np.concatenate([np.zeros(A.shape[:-1] + (left,)), A, np.zeros(A.shape[:-1] + (right,))], axis=-1)
    cs = np.concatenate([np.zeros(A.shape[:-1] + (1,)), np.cumsum(pad, axis=-1)], axis=-1)
    T = A.shape[-1]
    return (cs[..., n: n + T] - cs[..., 0:T]) / n


def torchgate_forward(x: np.ndarray, sr, xn: Optional[np.ndarray] = None, nonstationary=False,
                      n_std_thresh_stationary=1.5, n_thresh_nonstationary=1.3,
                      temp_coeff_nonstationary=0.1, n_movemean_nonstationary=20, prop_decrease=1.0,
                      n_fft=1024, win_length=None, hop_length=None, freq_mask_smooth_hz=500,
                      time_mask_smooth_ms=50, window: Optional[np.ndarray] = None,
                      filt: Optional[np.ndarray] = None, taps: Optional[dict] = None) -> np.ndarray:
    """TorchGate(...).forward(x, xn) (torchgate.py:200-264).  x: [B, L] -> [B, (L // hop) * hop]
    float64 (the reference casts back to x.dtype at :264; callers do that)."""
    assert x.ndim == 2
    W = n_fft if win_length is None else win_length
    H = W // 4 if hop_length is None else hop_length
    if x.shape[-1] < W * 2:
        raise Exception(f"x must be bigger than {W * 2}")
    if xn is not None and xn.shape[-1] < W * 2:
        raise Exception(f"xn must be bigger than {W * 2}")
    if window is None:
        window = hann_window_f32(W)
    smooth, nf, nt = smoothing_extents(sr, n_fft, H, freq_mask_smooth_hz, time_mask_smooth_ms)
    if smooth and filt is None:
        filt = smoothing_filter_f32(nf, nt)

    X = stft_center(x, n_fft, W, H, window)
    if nonstationary:
        A = np.abs(X)
        S = moving_mean_same(A, n_movemean_nonstationary)
        with np.errstate(divide="ignore", invalid="ignore"):
            ratio = (A - S) / S
            mask = 1.0 / (1.0 + np.exp(-(ratio - n_thresh_nonstationary) / temp_coeff_nonstationary))
        mask0 = mask
    else:
        X_db = amp_to_db(X)
        if xn is not None:
            xn2 = xn[None, :] if xn.ndim == 1 else xn
            XN_db = amp_to_db(stft_center(xn2, n_fft, W, H, window))
        else:
            XN_db = X_db
        mean = XN_db.mean(axis=-1)
        std = XN_db.std(axis=-1, ddof=1)                          # torch.std_mean is unbiased
        thresh = mean + std * n_std_thresh_stationary             # [Bn, F]
        mask0 = X_db > thresh[:, :, None]
        mask = mask0 * 1.0
        if taps is not None:
            taps.update(thresh=thresh, X_db=X_db)
    mask = prop_decrease * (mask * 1.0 - 1.0) + 1.0               # torchgate.py:241
    if smooth:
        mask = conv2d_same(mask, filt)
    if taps is not None:
        taps.update(X=X, mask0=mask0, mask=mask)
    return istft_center(X * mask, n_fft, W, H, window)
