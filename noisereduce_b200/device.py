"""Device-resident use of the operator: torch CUDA tensors in, torch CUDA tensors out.

PyTorch is plumbing here (device memory, streams, torch.distributed); the arithmetic is the C ABI's.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _cabi
from .spectralgate.base import SpectralGate


class DeviceGate:
    """A gate bound to the current CUDA device, taking [C, N] float32 / int16 / float64 CUDA tensors
    (read and written by the kernels in that dtype).  (With the CPU
    simulator library that tests inject via `lib=`, "device" pointers are host pointers.)"""

    def __init__(self, sr, stationary=True, prop_decrease=1.0, time_constant_s=2.0, freq_mask_smooth_hz=500,
                 time_mask_smooth_ms=50, thresh_n_mult_nonstationary=2, sigmoid_slope_nonstationary=10,
                 n_std_thresh_stationary=1.5, chunk_size=600000, padding=30000, n_fft=1024, win_length=None,
                 hop_length=None, clip_noise_stationary=True, lib=None, reserve_sms=0, workspace_limit_bytes=0.0,
                 path_flags=0):
        # reuse the reference-mirroring argument resolution (defaults, smoothing extents, errors)
        geo = SpectralGate(np.zeros(1, np.float32), sr, prop_decrease, chunk_size, padding, n_fft, win_length,
                           hop_length, time_constant_s, freq_mask_smooth_hz, time_mask_smooth_ms, None, False, 1)
        p = geo._gate_params()
        p.update(stationary=1 if stationary else 0, n_std_thresh=float(n_std_thresh_stationary),
                 clip_noise=1 if clip_noise_stationary else 0, time_constant_s=float(time_constant_s),
                 thresh_n_mult=float(thresh_n_mult_nonstationary), sigmoid_slope=float(sigmoid_slope_nonstationary),
                 reserve_sms=int(reserve_sms), workspace_limit_bytes=float(workspace_limit_bytes),
                 path_flags=int(path_flags))
        self.gate = _cabi.Gate(lib=lib, **p)
        self.stationary = stationary

    _NP = {torch.float32: np.float32, torch.int16: np.int16, torch.float64: np.float64}

    @classmethod
    def _check(cls, x: torch.Tensor):
        if not (x.dim() == 2 and x.dtype in cls._NP and x.stride(1) == 1):
            raise ValueError("expected a [C, N] float32 / int16 / float64 tensor with contiguous rows")

    def noise_stats(self, y_noise: torch.Tensor):
        self._check(y_noise)
        st = torch.cuda.current_stream().cuda_stream if y_noise.is_cuda else None
        self.gate.noise_stats_device(y_noise.data_ptr(), self._NP[y_noise.dtype], y_noise.shape[0], y_noise.shape[1],
                                     y_noise.stride(0), st)

    def run(self, x: torch.Tensor, out: torch.Tensor = None) -> torch.Tensor:
        self._check(x)
        if out is None:
            out = torch.empty_like(x)
        self._check(out)
        if out.dtype != x.dtype:
            raise ValueError("out must have x's dtype (the reference returns the input dtype)")
        st = torch.cuda.current_stream().cuda_stream if x.is_cuda else None
        self.gate.run_device(x.data_ptr(), out.data_ptr(), self._NP[x.dtype], x.shape[0], x.shape[1], x.stride(0),
                             out.stride(0), st)
        return out

    def run_chunks(self, x: torch.Tensor, slab: torch.Tensor, first: int, last: int) -> torch.Tensor:
        """Denoise only chunks [first, last] of the chunk grid (base.py:175-217) and write them densely into
        `slab` ([C, >= range length], contiguous rows): the library addresses the output through the virtual
        base `slab - first * chunk_size` (b200gate_set_range mode 1).  Returns the written view."""
        self._check(x)
        self._check(slab)
        cs = int(self.gate.params.chunk_size)
        C, N = x.shape
        if cs <= 0 or N <= cs:
            raise ValueError("run_chunks needs a chunked recording (N > chunk_size)")
        lo, hi = first * cs, min(N, (last + 1) * cs)
        if slab.dtype != x.dtype or slab.shape[0] != C or slab.shape[1] < hi - lo:
            raise ValueError("slab must be [C, >= range length] in x's dtype")
        st = torch.cuda.current_stream().cuda_stream if x.is_cuda else None
        self.gate.set_range(1, first, last)
        try:
            self.gate.run_device(x.data_ptr(), slab.data_ptr() - lo * x.element_size(), self._NP[x.dtype], C, N,
                                 x.stride(0), slab.stride(0), st)
        finally:
            self.gate.set_range(0)
        return slab[:, : hi - lo]
