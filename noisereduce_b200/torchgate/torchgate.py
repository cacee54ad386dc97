"""TorchGate -- drop-in for noisereduce/torchgate/torchgate.py:7 on libb200gate.

Same constructor signature, attributes, registered ``smoothing_filter`` buffer (so ``state_dict``
round-trips), ``forward(x, xn)`` contract and exceptions as the reference.  The arithmetic is not
torch ops: forward() hands the tensors' device pointers to the C ABI on the current CUDA stream.
CUDA tensors only -- there is no CPU path.
"""
from typing import Optional, Union

import numpy as np
import torch

from .. import _cabi


class TorchGate(torch.nn.Module):
    @torch.no_grad()
    def __init__(
        self,
        sr: int,
        nonstationary: bool = False,
        n_std_thresh_stationary: float = 1.5,
        n_thresh_nonstationary: float = 1.3,
        temp_coeff_nonstationary: float = 0.1,
        n_movemean_nonstationary: int = 20,
        prop_decrease: float = 1.0,
        n_fft: int = 1024,
        win_length: int = None,
        hop_length: int = None,
        freq_mask_smooth_hz: float = 500,
        time_mask_smooth_ms: float = 50,
    ):
        super().__init__()
        self.sr = sr
        self.nonstationary = nonstationary
        assert 0.0 <= prop_decrease <= 1.0                                     # torchgate.py:52
        self.prop_decrease = prop_decrease
        self.n_fft = n_fft
        self.win_length = self.n_fft if win_length is None else win_length      # torchgate.py:56-58
        self.hop_length = self.win_length // 4 if hop_length is None else hop_length
        self.n_std_thresh_stationary = n_std_thresh_stationary
        self.temp_coeff_nonstationary = temp_coeff_nonstationary
        self.n_movemean_nonstationary = n_movemean_nonstationary
        self.n_thresh_nonstationary = n_thresh_nonstationary
        self.freq_mask_smooth_hz = freq_mask_smooth_hz
        self.time_mask_smooth_ms = time_mask_smooth_ms
        self._extents = (0, 0)
        self.register_buffer("smoothing_filter", self._generate_mask_smoothing_filter())
        self._gate = None
        self._gate_lib = None

    @torch.no_grad()
    def _generate_mask_smoothing_filter(self) -> Union[torch.Tensor, None]:
        """torchgate.py:74-124: (1, 1, 2 nf + 1, 2 nt + 1) float32 buffer, or None when disabled.
        The kernels use only the extents (the taps are (n+1-|k|) / (n+1)^2)."""
        if self.freq_mask_smooth_hz is None and self.time_mask_smooth_ms is None:
            return None
        nf = 1 if self.freq_mask_smooth_hz is None else int(self.freq_mask_smooth_hz / (self.sr / (self.n_fft / 2)))
        if nf < 1:
            raise ValueError(f"freq_mask_smooth_hz needs to be at least {int((self.sr / (self.n_fft / 2)))} Hz")
        nt = 1 if self.time_mask_smooth_ms is None else int(self.time_mask_smooth_ms / ((self.hop_length / self.sr) * 1000))
        if nt < 1:
            raise ValueError(f"time_mask_smooth_ms needs to be at least {int((self.hop_length / self.sr) * 1000)} ms")
        if nt == 1 and nf == 1:
            return None
        self._extents = (nf, nt)

        def tri(n):
            k = torch.arange(-n, n + 1, dtype=torch.float32)
            return (n + 1 - k.abs()) / (n + 1)

        filt = torch.outer(tri(nf), tri(nt))[None, None]
        return filt / filt.sum()

    def _get_gate(self, lib=None):
        if self._gate is None or self._gate_lib is not lib:
            nf, nt = self._extents if self.smoothing_filter is not None else (0, 0)
            self._gate = _cabi.Gate(
                lib=lib, surface=_cabi.SURFACE_TORCH, stationary=0 if self.nonstationary else 1,
                n_fft=int(self.n_fft), win_length=int(self.win_length), hop_length=int(self.hop_length),
                n_grad_freq=nf, n_grad_time=nt, std_ddof=1, chunk_size=0, padding=0, sr=float(self.sr),
                prop_decrease=float(self.prop_decrease), n_std_thresh=float(self.n_std_thresh_stationary),
                top_db=40.0,                                                   # torchgate/utils.py:6
                n_movemean=int(self.n_movemean_nonstationary), thresh_n_mult=float(self.n_thresh_nonstationary),
                sigmoid_slope=1.0 / float(self.temp_coeff_nonstationary))
            # the reference's own window table (torchgate.py:231): torch.hann_window, float32
            self._gate.set_window(torch.hann_window(self.win_length).numpy())
            self._gate_lib = lib
        return self._gate

    def forward(self, x: torch.Tensor, xn: Optional[torch.Tensor] = None, _lib=None) -> torch.Tensor:
        assert x.ndim == 2                                                      # torchgate.py:214-220
        if x.shape[-1] < self.win_length * 2:
            raise Exception(f"x must be bigger than {self.win_length * 2}")
        assert xn is None or xn.ndim == 1 or xn.ndim == 2
        if xn is not None and xn.shape[-1] < self.win_length * 2:
            raise Exception(f"xn must be bigger than {self.win_length * 2}")
        if _lib is None and not x.is_cuda:
            raise RuntimeError("noisereduce_b200.TorchGate runs on CUDA tensors only (no CPU fallback)")
        if torch.is_grad_enabled() and x.requires_grad:
            # The reference builds its masks under no_grad but lets the gradient flow through
            # stft -> (* mask) -> istft (torchgate.py:223-262).  The CUDA path has no backward kernel: refuse loudly
            # rather than hand back a tensor that silently stops the gradient.
            raise NotImplementedError(
                "noisereduce_b200.TorchGate has no backward pass; call it under torch.no_grad() or on x.detach()")
        gate = self._get_gate(_lib)
        stream = torch.cuda.current_stream().cuda_stream if x.is_cuda else None
        xf = x.detach().to(torch.float32).contiguous()
        if xn is not None and not self.nonstationary:
            xnf = xn.detach().to(device=x.device, dtype=torch.float32)
            xnf = (xnf[None, :] if xnf.ndim == 1 else xnf).contiguous()
            gate.torch_set_noise(xnf.data_ptr(), xnf.shape[0], xnf.shape[1], xnf.stride(0), True, stream)
        else:
            gate.torch_set_noise(None, 0, 0, 0, True, stream)
        B, L = xf.shape
        Lo = (L // self.hop_length) * self.hop_length + (self.n_fft & 1)     # torch.istft(center=True) length
        y = torch.empty((B, Lo), dtype=torch.float32, device=x.device)
        gate.run_device(xf.data_ptr(), y.data_ptr(), np.float32, B, L, xf.stride(0), y.stride(0), stream)
        return y.to(dtype=x.dtype)                                              # torchgate.py:264
