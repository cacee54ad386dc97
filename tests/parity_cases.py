"""Parity checks shared by the CPU-simulator tests (tests/test_cusim_parity.py, small sizes) and the
real-GPU tests (tests/test_gpu_parity.py).  Every check compares the library behind `lib` against
oracle/ on the same seeded input, stage by stage:

  * binary mask decisions: BIT-EXACT against the float64 oracle (thresholds injected so both sides
    threshold against the very same numbers);
  * FP32 STFT, smoothed mask, output: within the float32 tolerances written below;
  * the library's own noise statistics: thresholds within THRESH_TOL_DB of the oracle's.
"""
import numpy as np

from noisereduce_b200 import _cabi
from oracle import spectral_gate_oracle as O

OUT_TOL = 1e-4            # rel-inf on the waveform: max|y - y_ref| / max|y_ref|  (BASELINE.json target)
OUT_TOL_TIGHT = 2e-6      # what the FP32 path actually achieves when the mask decisions agree
SPEC_TOL = 2e-6           # FP32 STFT vs float64, relative to the largest bin
MASK_TOL = 1e-6           # smoothed mask, absolute (values in [0, 1])
MASK_TOL_NONSTAT = 2e-5   # sigmoid mask: slope 10 amplifies the FP32 error of (|X| - S) / S
THRESH_TOL_DB = 1e-3      # our FP64 noise statistics vs the reference's float32-STFT statistics


def relinf(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def gate_params(cfg: O.GateConfig, **extra):
    N, W, H = cfg.resolve()
    smooth, nf, nt = O.smoothing_extents(cfg.sr, N, H, cfg.freq_mask_smooth_hz, cfg.time_mask_smooth_ms)
    p = dict(
        surface=_cabi.SURFACE_NUMPY, stationary=1 if cfg.stationary else 0, n_fft=N, win_length=W, hop_length=H,
        n_grad_freq=nf if smooth else 0, n_grad_time=nt if smooth else 0,
        chunk_size=cfg.chunk_size if cfg.chunk_size is not None else 0, padding=cfg.padding, sr=float(cfg.sr),
        prop_decrease=cfg.prop_decrease, n_std_thresh=cfg.n_std_thresh_stationary, top_db=80.0, std_ddof=0,
        clip_noise=1 if cfg.clip_noise_stationary else 0, time_constant_s=cfg.time_constant_s,
        thresh_n_mult=float(cfg.thresh_n_mult_nonstationary), sigmoid_slope=float(cfg.sigmoid_slope_nonstationary),
    )
    p.update(extra)
    return p


def check_stationary(lib, y, cfg: O.GateConfig, tap_unit=(0, 0), y_noise=None, inject_thresh=True, **extra):
    """Returns a dict of measured errors (asserting is left to the caller so reports can show them)."""
    y2d = y if y.ndim == 2 else y[None, :]
    taps = {tap_unit: O.Taps()}
    info = {}
    ref = O.reduce_noise(y2d, cfg.sr, y_noise=y_noise, cfg=cfg, unit_taps=taps, info=info, return_float64=True)
    gate = _cabi.Gate(lib=lib, **gate_params(cfg, **extra))
    noise = y2d if y_noise is None else (y_noise if y_noise.ndim == 2 else y_noise[None, :])
    gate.noise_stats_host(noise)
    res = dict(thresh_err_db=float(np.abs(gate.noise_threshold() - info["thresh"]).max()))
    if inject_thresh:
        gate.set_noise_threshold(info["thresh"])
    gate.debug_select_unit(*tap_unit)
    out = gate.run_host(y2d)
    res["stats"] = gate.stats()
    d = gate.debug_read()
    tp = taps[tap_unit]
    res["T"] = tp.X.shape[1]
    # the single-pass kernel only transforms the frames its output needs (chunk centre + halos): compare the
    # analysis taps on the frames the kernel touched (all of them on the two-pass path)
    seen = (d["X"] != 0).any(axis=0) | ~(tp.X != 0).any(axis=0)
    res["analysis_frames_checked"] = int(seen.sum())
    res["spec_err"] = float(np.abs(d["X"] - tp.X)[:, seen].max() / np.abs(tp.X).max())
    res["mask0_mismatch"] = int((d["mask0"] != tp.mask0)[:, seen].sum())
    res["mask0_on_frac"] = float(tp.mask0.mean())
    touched = d["mask"].any(axis=0) | ~tp.mask.any(axis=0)
    res["mask_frames_checked"] = int(touched.sum())
    res["mask_err"] = float(np.abs(d["mask"] - tp.mask)[:, touched].max()) if touched.any() else 0.0
    out_ref = O.cast_like_reference(ref, y2d.dtype)
    res["out_dtype_ok"] = out.dtype == y2d.dtype
    if np.issubdtype(y2d.dtype, np.integer):
        res["out_max_lsb"] = int(np.abs(out.astype(np.int64) - out_ref.astype(np.int64)).max())
    res["out_relinf"] = relinf(out, ref)
    res["out"] = out
    res["ref"] = ref
    gate.close()
    return res


def check_nonstationary(lib, y, cfg: O.GateConfig, tap_unit=(0, 0), **extra):
    y2d = y if y.ndim == 2 else y[None, :]
    taps = {tap_unit: O.Taps()}
    ref = O.reduce_noise(y2d, cfg.sr, cfg=cfg, unit_taps=taps, return_float64=True)
    gate = _cabi.Gate(lib=lib, **gate_params(cfg, **extra))
    gate.debug_select_unit(*tap_unit)
    out = gate.run_host(y2d)
    d = gate.debug_read()
    tp = taps[tap_unit]
    res = dict(stats=gate.stats(), T=tp.X.shape[1])
    res["spec_err"] = float(np.abs(d["X"] - tp.X).max() / np.abs(tp.X).max())
    touched = d["mask"].any(axis=0)
    res["mask_frames_checked"] = int(touched.sum())
    res["mask_err"] = float(np.abs(d["mask"] - tp.mask)[:, touched].max()) if touched.any() else 0.0
    res["out_dtype_ok"] = out.dtype == y2d.dtype
    res["out_relinf"] = relinf(out, ref)
    gate.close()
    return res


def reference_test_suite_scenarios(fish_int16, sr, n=None):
    """The four numpy-path scenarios of the reference's test_reduction.py (:6-56), on its own asset: float64
    `fish + noise * 10`, stationary with / without a 2-second noise clip, non-stationary, non-stationary in
    30000-sample chunks.  Yields (name, y, kwargs)."""
    from tests.synth_host import band_noise
    data = fish_int16[: n] if n else fish_int16
    noise = band_noise(len(data), sr) * 10
    y = data + noise                                           # int16 + float64 -> float64, as in the reference's tests
    yield "stationary_with_noise_clip", y, dict(stationary=True, y_noise=noise[: sr * 2])
    yield "stationary_without_noise_clip", y, dict(stationary=True)
    yield "nonstationary", y, dict(stationary=False)
    yield "nonstationary_batches", y, dict(stationary=False, chunk_size=30000)
