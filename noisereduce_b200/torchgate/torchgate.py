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

    def _params(self):
        return (int(self.sr), bool(self.nonstationary), float(self.n_std_thresh_stationary), float(self.n_thresh_nonstationary),
                float(self.temp_coeff_nonstationary), int(self.n_movemean_nonstationary), float(self.prop_decrease),
                int(self.n_fft), int(self.win_length), int(self.hop_length))

    def _get_gate(self, lib, device, f64):
        """One library handle per (library, device, sample type, parameter values): the handle's tables live on the
        device that was current when it was created, and the reference reads its attributes on every forward."""
        key = (id(lib), str(device), bool(f64), self._params())
        if self._gate is None:
            self._gate = {}
        gate = self._gate.get(key)
        if gate is None:
            nf, nt = self._extents if self.smoothing_filter is not None else (0, 0)
            gate = _cabi.Gate(
                lib=lib, surface=_cabi.SURFACE_TORCH, stationary=0 if self.nonstationary else 1,
                n_fft=int(self.n_fft), win_length=int(self.win_length), hop_length=int(self.hop_length),
                n_grad_freq=nf, n_grad_time=nt, std_ddof=1, chunk_size=0, padding=0, sr=float(self.sr),
                prop_decrease=float(self.prop_decrease), n_std_thresh=float(self.n_std_thresh_stationary),
                top_db=40.0,                                                   # torchgate/utils.py:6
                n_movemean=int(self.n_movemean_nonstationary), thresh_n_mult=float(self.n_thresh_nonstationary),
                sigmoid_slope=1.0 / float(self.temp_coeff_nonstationary),
                path_flags=4 if f64 else 0)           # float64 input: the float64 general-geometry family (complex128 like the reference)
            # the reference's own window table (torchgate.py:231): torch.hann_window, float32
            gate.set_window(torch.hann_window(self.win_length).numpy())
            self._gate[key] = gate
        return gate

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
            # torchgate.py:223-262: the masks are built under no_grad, the gradient flows through stft -> * mask -> istft
            return _GateFunction.apply(x, self, xn, _lib)
        return self._forward_impl(x, xn, _lib)[0]

    def _tuned(self):
        return self.n_fft == 1024 and self.win_length == 1024 and self.hop_length == 256

    def _forward_impl(self, x, xn, _lib):
        f64 = x.dtype == torch.float64
        work_t, np_t = (torch.float64, np.float64) if f64 else (torch.float32, np.float32)
        ctx = torch.cuda.device(x.device) if x.is_cuda else _NullCtx()
        with ctx:
            gate = self._get_gate(_lib, x.device, f64)
            stream = torch.cuda.current_stream(x.device).cuda_stream if x.is_cuda else None
            xf = x.detach().to(work_t).contiguous()
            if xn is not None and not self.nonstationary:
                xnf = xn.detach().to(device=x.device, dtype=work_t)
                xnf = (xnf[None, :] if xnf.ndim == 1 else xnf).contiguous()
                gate.torch_set_noise(xnf.data_ptr(), xnf.shape[0], xnf.shape[1], xnf.stride(0), True, stream, dtype=np_t)
            else:
                gate.torch_set_noise(None, 0, 0, 0, True, stream)
            B, L = xf.shape
            Lo = (L // self.hop_length) * self.hop_length + (self.n_fft & 1)     # torch.istft(center=True) length
            y = torch.empty((B, Lo), dtype=work_t, device=x.device)
            gate.run_device(xf.data_ptr(), y.data_ptr(), np_t, B, L, xf.stride(0), y.stride(0), stream)
            gate._fwd_serial = getattr(gate, "_fwd_serial", 0) + 1
        return y.to(dtype=x.dtype), gate                                        # torchgate.py:264

    def _ola_envelope(self, L, device):
        """sum_t w^2 of torch.istft(center=True) for T = 1 + L // hop frames, on the output's Lo samples."""
        N, H = self.n_fft, self.hop_length
        T = 1 + L // H
        w2 = torch.hann_window(self.win_length, device=device, dtype=torch.float32) ** 2
        env = torch.zeros(N + H * (T - 1), device=device, dtype=torch.float32)
        idx = (torch.arange(T, device=device) * H)[:, None] + torch.arange(N, device=device)[None, :]
        env.index_add_(0, idx.reshape(-1), w2.repeat(T))
        return env[N // 2: N // 2 + (L // H) * H]

    def _adjoint(self, g, gate, serial, L, _lib):
        """Gradient of forward with respect to x (masks fixed): the gate is  y = OLA(w B(w frame(x))) / env  with a
        symmetric B (a real mask between rfft and irfft), so  dL/dx = env * gate(g / env)  with the SAME masks --
        one more analysis + synthesis through b200gate_torch_apply_masks.  Input samples beyond the output length
        ((L // hop) * hop) get a zero gradient."""
        if getattr(gate, "_fwd_serial", 0) != serial:
            raise RuntimeError("TorchGate.backward: another forward ran on this module since the forward being "
                               "differentiated (its masks were overwritten); call backward first or use a second module")
        B, Lo = g.shape
        env = self._ola_envelope(L, g.device)
        z = torch.zeros((B, L), dtype=torch.float32, device=g.device)
        z[:, :Lo] = g.to(torch.float32) / env
        out = torch.empty((B, Lo), dtype=torch.float32, device=g.device)
        ctx = torch.cuda.device(g.device) if g.is_cuda else _NullCtx()
        with ctx:
            stream = torch.cuda.current_stream(g.device).cuda_stream if g.is_cuda else None
            gate.torch_apply_masks_device(z.data_ptr(), out.data_ptr(), np.float32, B, L, z.stride(0), out.stride(0), stream)
        grad = torch.zeros((B, L), dtype=torch.float32, device=g.device)
        grad[:, :Lo] = out * env
        return grad


class _GateFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, module, xn, lib):
        if not module._tuned() or x.dtype == torch.float64:
            raise NotImplementedError("noisereduce_b200.TorchGate: the backward pass exists for float32 / float16 input at the "
                                      "default STFT geometry (n_fft 1024, hop 256); use torch.no_grad() otherwise")
        y, gate = module._forward_impl(x, xn, lib)
        ctx.module, ctx.gate, ctx.lib = module, gate, lib
        ctx.serial, ctx.L, ctx.dtype = gate._fwd_serial, x.shape[-1], x.dtype
        return y

    @staticmethod
    def backward(ctx, g):
        grad = ctx.module._adjoint(g.contiguous(), ctx.gate, ctx.serial, ctx.L, ctx.lib)
        return grad.to(ctx.dtype), None, None, None


class _NullCtx:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False
