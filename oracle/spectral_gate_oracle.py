"""CPU oracle for the spectral-gating hot path (TEST INFRASTRUCTURE -- see oracle/__init__.py).

Every function cites the reference file:line (relative to /root/reference) or the SciPy 1.18.1
routine whose published algorithm it restates.  All chunk arithmetic is float64, exactly as the
reference promotes every chunk to float64 (noisereduce/spectralgate/base.py:140).

Notation: N = n_fft, W = win_length, H = hop_length, F = N//2 + 1,
w = periodic Hann of length W, T = Lp//H + 1 frames for a padded chunk of Lp samples.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np
import scipy.fft as _sfft

EPS64 = float(np.finfo(np.float64).eps)  # spectralgate/utils.py:11 default eps


# --------------------------------------------------------------------------------------
# DFT primitive.  scipy.signal.stft/istft call scipy.fft.rfft/irfft (pocketfft); we keep that one
# primitive and pin it against `dft_matrix_rfft` (plain O(N^2) definition) in the tests.
# --------------------------------------------------------------------------------------
def rfft(frames: np.ndarray, n: int) -> np.ndarray:
    return _sfft.rfft(frames, n=n, axis=-1)


def irfft(spec: np.ndarray, n: int) -> np.ndarray:
    return _sfft.irfft(spec, n=n, axis=-1)


def dft_matrix_rfft(frames: np.ndarray, n: int) -> np.ndarray:
    """Definition of the one-sided DFT, X[f] = sum_k x[k] exp(-2 pi i f k / n) (float64)."""
    frames = np.asarray(frames, dtype=np.float64)
    k = np.arange(frames.shape[-1])
    f = np.arange(n // 2 + 1)
    ang = -2.0 * np.pi * ((f[:, None] * k[None, :]) % n) / n
    return frames @ np.exp(1j * ang).T


# --------------------------------------------------------------------------------------
# Window / smoothing filter / geometry
# --------------------------------------------------------------------------------------
def hann_periodic(W: int) -> np.ndarray:
    """scipy.signal.get_window('hann_periodic', W) (== 'hann', fftbins=True): 0.5 - 0.5 cos(2 pi k / W)."""
    k = np.arange(W, dtype=np.float64)
    return 0.5 - 0.5 * np.cos(2.0 * np.pi * k / W)


def triangle_taps(n: int) -> np.ndarray:
    """One axis of _smoothing_filter (base.py:14-28): v[k] = (n + 1 - |k|) / (n + 1), |k| <= n."""
    k = np.arange(-n, n + 1, dtype=np.float64)
    return (n + 1 - np.abs(k)) / (n + 1)


def smoothing_filter(n_grad_freq: int, n_grad_time: int) -> np.ndarray:
    """base.py:7-29: outer product of the two triangles, normalised to unit sum."""
    filt = np.outer(triangle_taps(n_grad_freq), triangle_taps(n_grad_time))
    return filt / np.sum(filt)


def smoothing_extents(sr, n_fft, hop_length, freq_mask_smooth_hz, time_mask_smooth_ms) -> Tuple[bool, int, int]:
    """base.py:92-128.  Returns (smooth_mask, n_grad_freq, n_grad_time); raises the same ValueErrors."""
    if freq_mask_smooth_hz is None and time_mask_smooth_ms is None:
        return False, 1, 1
    if freq_mask_smooth_hz is None:
        n_grad_freq = 1
    else:
        n_grad_freq = int(freq_mask_smooth_hz / (sr / (n_fft / 2)))
        if n_grad_freq < 1:
            raise ValueError(
                "freq_mask_smooth_hz needs to be at least {}Hz".format(int((sr / (n_fft / 2))))
            )
    if time_mask_smooth_ms is None:
        n_grad_time = 1
    else:
        n_grad_time = int(time_mask_smooth_ms / ((hop_length / sr) * 1000))
        if n_grad_time < 1:
            raise ValueError(
                "time_mask_smooth_ms needs to be at least {}ms".format(int((hop_length / sr) * 1000))
            )
    if n_grad_time == 1 and n_grad_freq == 1:
        return False, 1, 1
    return True, n_grad_freq, n_grad_time


def chunk_table(n_frames: int, chunk_size: Optional[int], padding: int,
                start_frame: Optional[int] = None, end_frame: Optional[int] = None) -> List[Tuple[int, int, int, int]]:
    """base.py:167-226 (get_traces; start_frame / end_frame default to the whole recording).

    Returns a list of (i1, i2, out_lo, out_hi): the padded span [i1, i2) read by _read_chunk
    (base.py:130-142, zeros outside [0, n_frames)) and the output span [out_lo, out_hi) its centre
    fills.  Chunked branch (base.py:175-217): chunks int(start/cs) .. int((end-1)/cs) of the grid
    anchored at sample 0, trimmed to [start, end).  Otherwise (base.py:222) ONE padded chunk
    covering [0, end_frame) -- start_frame is ignored there, as in the reference.
    """
    start = 0 if start_frame is None else start_frame
    end = n_frames if end_frame is None else end_frame
    if chunk_size is not None and end - start > chunk_size:
        tab = []
        for ich in range(int(start / chunk_size), int((end - 1) / chunk_size) + 1):
            s, e = ich * chunk_size, (ich + 1) * chunk_size
            tab.append((s - padding, e + padding, max(s, start), min(e, end)))
        return tab
    return [(-padding, end + padding, 0, end)]


def traces_span(n_frames: int, chunk_size: Optional[int], start_frame: Optional[int] = None,
                end_frame: Optional[int] = None) -> Tuple[int, int]:
    """The sample span get_traces returns (base.py:175 vs :222)."""
    start = 0 if start_frame is None else start_frame
    end = n_frames if end_frame is None else end_frame
    if chunk_size is not None and end - start > chunk_size:
        return start, end
    return 0, end


def read_chunk(y2d: np.ndarray, i1: int, i2: int) -> np.ndarray:
    """base.py:130-142: float64 zeros [C, i2-i1] with the valid span copied in."""
    n = y2d.shape[1]
    lo, hi = max(i1, 0), min(i2, n)
    chunk = np.zeros((y2d.shape[0], i2 - i1))
    if hi > lo:
        chunk[:, lo - i1: hi - i1] = y2d[:, lo:hi]
    return chunk


# --------------------------------------------------------------------------------------
# STFT / iSTFT exactly as the reference calls them
# --------------------------------------------------------------------------------------
def stft(x: np.ndarray, n_fft: int, win_length: int, hop_length: int) -> np.ndarray:
    """scipy.signal.stft(x, nfft=N, nperseg=W, noverlap=W-H, padded=False) as called at
    stationary.py:67-73, :87-93 and nonstationary.py:51-57 (defaults window='hann' periodic,
    boundary='zeros', scaling='spectrum', one-sided).  scipy:_spectral_py.py:_spectral_helper.

    x: 1-D real.  Returns X[F, T].  Computation dtype follows x (float32 -> complex64 with the
    window rounded to float32, float64 -> complex128), as _spectral_helper does.
    """
    x = np.asarray(x)
    W, H, N = int(win_length), int(hop_length), int(n_fft)
    real_dt = np.float32 if x.dtype == np.float32 else np.float64
    x = x.astype(real_dt, copy=False)
    w = hann_periodic(W).astype(real_dt)
    xe = np.concatenate([np.zeros(W // 2, real_dt), x, np.zeros(W // 2, real_dt)])
    if xe.shape[0] < W:
        raise ValueError("signal shorter than one window")
    T = (xe.shape[0] - W) // H + 1
    idx = (np.arange(T) * H)[:, None] + np.arange(W)[None, :]
    frames = w[None, :] * xe[idx]                      # [T, W]   window multiply in x's precision
    X = rfft(frames, N)                                # zero-pads W -> N at the END
    scale = real_dt(1.0) / w.sum(dtype=real_dt)        # sqrt(1 / win.sum()**2)
    X = X * scale
    return np.ascontiguousarray(X.T)                   # [F, T]


def istft(X: np.ndarray, n_fft: int, win_length: int, hop_length: int) -> np.ndarray:
    """scipy.signal.istft(X, nfft=N, nperseg=W, noverlap=W-H) as called at stationary.py:120-125 /
    nonstationary.py:90-95 (boundary=True, scaling='spectrum').  scipy:_spectral_py.py:istft.
    X: [F, T] complex128 -> real signal of length (T-1)*H.
    """
    W, H, N = int(win_length), int(hop_length), int(n_fft)
    T = X.shape[1]
    w = hann_periodic(W)
    seg = irfft(np.ascontiguousarray(X.T), N)[:, :W] * w.sum()        # [T, W]
    L = W + (T - 1) * H
    out = np.zeros(L)
    norm = np.zeros(L)
    for t in range(T):
        out[t * H: t * H + W] += seg[t] * w
        norm[t * H: t * H + W] += w * w
    out = out[W // 2: L - W // 2]
    norm = norm[W // 2: L - W // 2]
    return out / np.where(norm > 1e-10, norm, 1.0)


# --------------------------------------------------------------------------------------
# dB, thresholds, masks, smoothing
# --------------------------------------------------------------------------------------
def amp_to_db(X: np.ndarray, top_db: float = 80.0, eps: float = EPS64) -> np.ndarray:
    """spectralgate/utils.py:11-16.  np.abs keeps X's precision; `+ np.float64 eps` promotes to
    float64 (NEP 50), so the log is always float64.  Floor = row max over the LAST axis - top_db."""
    x_db = 20 * np.log10(np.abs(X) + np.float64(eps))
    return np.maximum(x_db, np.max(x_db, axis=-1, keepdims=True) - top_db)


def collapse_noise(y_noise2d: np.ndarray, chunk_size: Optional[int], clip: bool) -> np.ndarray:
    """stationary.py:61-64: channel mean in the input dtype (float32 stays float32; integer input
    promotes to float64), then clip to the first chunk_size samples."""
    yn = np.mean(y_noise2d, axis=0)
    if clip:
        yn = yn[:chunk_size]
    return yn


def stationary_threshold(y_noise1d: np.ndarray, n_fft, win_length, hop_length, n_std: float):
    """stationary.py:67-81.  Returns (thresh[F], mean[F], std[F], noise_db[F, Tn])."""
    Xn = stft(y_noise1d, n_fft, win_length, hop_length)
    db = amp_to_db(Xn)
    mean = np.mean(db, axis=1)
    std = np.std(db, axis=1)          # population std (ddof 0)
    return mean + std * n_std, mean, std, db


def conv2d_same_zero(mask: np.ndarray, filt: np.ndarray) -> np.ndarray:
    """scipy.signal.fftconvolve(mask, filt, mode='same') (stationary.py:114, nonstationary.py:80),
    restated as the direct zero-padded linear convolution it approximates, using separability of
    the filter (outer product of two symmetric triangles, base.py:14-28)."""
    kf, kt = filt.shape
    nf, nt = kf // 2, kt // 2
    # separable factors: filt = outer(vf, vt) / (sum vf * sum vt)
    vf = filt[:, nt] / filt[nf, nt]
    vt = filt[nf, :] / filt[nf, nt]
    norm = filt[nf, nt]
    F, T = mask.shape
    tmp = np.zeros((F, T))
    padded = np.zeros((F, T + 2 * nt))
    padded[:, nt: nt + T] = mask
    for b in range(kt):
        tmp += vt[b] * padded[:, 2 * nt - b: 2 * nt - b + T]
    out = np.zeros((F, T))
    padded = np.zeros((F + 2 * nf, T))
    padded[nf: nf + F, :] = tmp
    for a in range(kf):
        out += vf[a] * padded[2 * nf - a: 2 * nf - a + F, :]
    return out * norm


def filtfilt_onepole(A: np.ndarray, b: float) -> np.ndarray:
    """scipy.signal.filtfilt([b], [1, b-1], A, axis=-1, padtype=None) (nonstationary.py:115):
    forward sweep s[n] = b x[n] + (1-b) s[n-1] started at s[-1] = x[0] (lfilter_zi steady state),
    then the same sweep backwards over the forward output, started at its last value."""
    F, T = A.shape
    fwd = np.empty((F, T))
    s = A[:, 0].copy()
    for n in range(T):
        s = b * A[:, n] + (1.0 - b) * s
        fwd[:, n] = s
    out = np.empty((F, T))
    s = fwd[:, T - 1].copy()
    for n in range(T - 1, -1, -1):
        s = b * fwd[:, n] + (1.0 - b) * s
        out[:, n] = s
    return out


def iir_coefficient(time_constant_s: float, sr, hop_length) -> float:
    """nonstationary.py:109-114."""
    t_frames = time_constant_s * sr / float(hop_length)
    return float((np.sqrt(1 + 4 * t_frames ** 2) - 1) / (2 * t_frames ** 2))


def sigmoid(x, shift, mult):
    """spectralgate/utils.py:4-8."""
    return 1 / (1 + np.exp(-(x + shift) * mult))


# --------------------------------------------------------------------------------------
# Per-(chunk, channel) gate with stage taps
# --------------------------------------------------------------------------------------
@dataclass
class GateConfig:
    sr: float
    stationary: bool = False
    prop_decrease: float = 1.0
    time_constant_s: float = 2.0
    freq_mask_smooth_hz: Optional[float] = 500
    time_mask_smooth_ms: Optional[float] = 50
    thresh_n_mult_nonstationary: float = 2
    sigmoid_slope_nonstationary: float = 10
    n_std_thresh_stationary: float = 1.5
    chunk_size: Optional[int] = 600000
    padding: int = 30000
    n_fft: int = 1024
    win_length: Optional[int] = None
    hop_length: Optional[int] = None
    clip_noise_stationary: bool = True

    def resolve(self):
        """base.py:79-86 defaults."""
        W = self.n_fft if self.win_length is None else self.win_length
        H = W // 4 if self.hop_length is None else self.hop_length
        return self.n_fft, W, H


@dataclass
class Taps:
    """Stage outputs of one (chunk, channel) unit, for stage-by-stage parity tests."""
    X: Optional[np.ndarray] = None          # [F, T] complex128
    db: Optional[np.ndarray] = None         # stationary: clamped dB
    mask0: Optional[np.ndarray] = None      # stationary: bool [F,T]; non-stationary: sigmoid mask
    mask: Optional[np.ndarray] = None       # final multiplicative mask [F, T]
    smooth_floor: Optional[np.ndarray] = None  # non-stationary: filtfilt output
    y: Optional[np.ndarray] = None          # [Lp] filtered padded chunk


def gate_stationary_unit(x: np.ndarray, thresh: np.ndarray, cfg: GateConfig, filt, taps: Optional[Taps] = None):
    """stationary.py:83-127 for one channel of one padded chunk."""
    N, W, H = cfg.resolve()
    X = stft(x, N, W, H)
    db = amp_to_db(X)
    mask0 = db > thresh[:, None]                                  # stationary.py:99-106
    p = cfg.prop_decrease
    mask = mask0 * p + np.ones(mask0.shape) * (1.0 - p)           # stationary.py:108-110
    if filt is not None:
        mask = conv2d_same_zero(mask, filt)                       # stationary.py:112-114
    y = np.zeros(x.shape[0])
    sig = istft(X * mask, N, W, H)                                # stationary.py:117-125
    y[: len(sig)] = sig                                           # stationary.py:126
    if taps is not None:
        taps.X, taps.db, taps.mask0, taps.mask, taps.y = X, db, mask0, mask, y
    return y


def gate_nonstationary_unit(x: np.ndarray, cfg: GateConfig, filt, taps: Optional[Taps] = None):
    """nonstationary.py:47-97 for one channel of one padded chunk."""
    N, W, H = cfg.resolve()
    X = stft(x, N, W, H)
    A = np.abs(X)
    b = iir_coefficient(cfg.time_constant_s, cfg.sr, H)
    S = filtfilt_onepole(A, b)                                    # nonstationary.py:62-67
    with np.errstate(divide="ignore", invalid="ignore"):
        ratio = (A - S) / S                                       # nonstationary.py:70
        mask0 = sigmoid(ratio, -cfg.thresh_n_mult_nonstationary, cfg.sigmoid_slope_nonstationary)
    mask = mask0
    if filt is not None:
        mask = conv2d_same_zero(mask, filt)                       # nonstationary.py:78-80
    p = cfg.prop_decrease
    mask = mask * p + np.ones(mask.shape) * (1.0 - p)             # nonstationary.py:82-84
    y = np.zeros(x.shape[0])
    sig = istft(X * mask, N, W, H)
    y[: len(sig)] = sig
    if taps is not None:
        taps.X, taps.mask0, taps.mask, taps.smooth_floor, taps.y = X, mask0, mask, S, y
    return y


# --------------------------------------------------------------------------------------
# Whole-signal driver == reduce_noise(..., use_torch=False)
# --------------------------------------------------------------------------------------
def _as_2d(y):
    """base.py:54-62."""
    y = np.array(y)
    if y.ndim == 1:
        return y[None, :], True
    if y.ndim > 2:
        raise ValueError("Waveform must be in shape (# frames, # channels)")
    return y, False


def cast_like_reference(out64: np.ndarray, dtype) -> np.ndarray:
    """base.py:184/:164 (memmap assignment) and :218-226: plain numpy float64 -> dtype cast
    (C-style truncation toward zero for integer dtypes)."""
    with np.errstate(invalid="ignore"):
        return out64.astype(dtype)


def reduce_noise(y, sr, y_noise=None, cfg: Optional[GateConfig] = None, return_float64=False,
                 unit_taps: Optional[Dict[Tuple[int, int], Taps]] = None, thresh_override=None,
                 info: Optional[dict] = None, start_frame: Optional[int] = None, end_frame: Optional[int] = None):
    """noisereduce/noisereduce.py:13-185 with use_torch=False, as a loop over chunk_table x channels.

    unit_taps: optional dict keyed (chunk_index, channel) -> Taps to be filled.
    start_frame / end_frame: SpectralGate.get_traces(start_frame, end_frame) (base.py:167-226) instead of
    the whole recording.
    """
    cfg = cfg or GateConfig(sr=sr)
    y2d, flat = _as_2d(y)
    dtype = y2d.dtype
    C, n = y2d.shape
    N, W, H = cfg.resolve()
    smooth, nf, nt = smoothing_extents(sr, N, H, cfg.freq_mask_smooth_hz, cfg.time_mask_smooth_ms)
    filt = smoothing_filter(nf, nt) if smooth else None

    thresh = None
    if cfg.stationary:
        if thresh_override is not None:
            thresh = np.asarray(thresh_override, dtype=np.float64)
        else:
            if y_noise is None:
                yn2d = y2d                                         # stationary.py:47-48
            else:
                yn2d, _ = _as_2d(y_noise)
            yn = collapse_noise(yn2d, cfg.chunk_size, cfg.clip_noise_stationary)
            thresh, mean, std, _ = stationary_threshold(yn, N, W, H, cfg.n_std_thresh_stationary)
            if info is not None:
                info.update(noise_mean=mean, noise_std=std)
        if info is not None:
            info.update(thresh=thresh)
    if info is not None:
        info.update(n_grad_freq=nf, n_grad_time=nt, smooth=smooth, filt=filt)

    out = np.zeros((C, n))
    for ich, (i1, i2, lo, hi) in enumerate(chunk_table(n, cfg.chunk_size, cfg.padding, start_frame, end_frame)):
        chunk = read_chunk(y2d, i1, i2)
        for c in range(C):
            taps = None
            if unit_taps is not None and (ich, c) in unit_taps:
                taps = unit_taps[(ich, c)]
            if cfg.stationary:
                yc = gate_stationary_unit(chunk[c], thresh, cfg, filt, taps)
            else:
                yc = gate_nonstationary_unit(chunk[c], cfg, filt, taps)
            out[c, lo:hi] = yc[lo - i1: hi - i1]                   # base.py:150, :164
    s0, s1 = traces_span(n, cfg.chunk_size, start_frame, end_frame)
    out = out[:, s0:s1]
    if return_float64:
        return out[0] if flat else out
    res = cast_like_reference(out, dtype)
    return res.flatten() if flat else res
