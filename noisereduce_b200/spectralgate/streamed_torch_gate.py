"""Mirror of noisereduce/spectralgate/streamed_torch_gate.py:7 (StreamedTorchGate): reduce_noise(use_torch=True).

The reference's torch route is NOT the numpy algorithm on a GPU: every padded chunk goes through TorchGate
(streamed_torch_gate.py:81-88), i.e. torch.stft(center=True) framing, per-chunk self statistics with top_db = 40 and an
unbiased std, the blend applied before the smoothing, and for the non-stationary gate a moving mean of
int(time_constant_s / hop * sr) frames instead of the filtfilt follower, with temp_coeff = 1 / sigmoid_slope and
n_thresh = thresh_n_mult_nonstationary (streamed_torch_gate.py:66-79).  This class keeps exactly that mapping on top of
noisereduce_b200.TorchGate (the CUDA kernels of the torch surface) and the reference's chunk loop (base.py:167-226).
Differences: chunks are computed in float32 (the reference feeds TorchGate the float64 chunk of base.py:140), and
there is no CPU path -- `device` must be a CUDA device.
"""
import numpy as np
import torch

from ..torchgate import TorchGate
from .base import SpectralGate


class StreamedTorchGate(SpectralGate):
    def __init__(
        self,
        y,
        sr,
        stationary=False,
        y_noise=None,
        prop_decrease=1.0,
        time_constant_s=2.0,
        freq_mask_smooth_hz=500,
        time_mask_smooth_ms=50,
        thresh_n_mult_nonstationary=2,
        sigmoid_slope_nonstationary=10,
        n_std_thresh_stationary=1.5,
        tmp_folder=None,
        chunk_size=600000,
        padding=30000,
        n_fft=1024,
        win_length=None,
        hop_length=None,
        clip_noise_stationary=True,
        use_tqdm=False,
        n_jobs=1,
        device="cuda",
        _lib=None,
    ):
        super().__init__(
            y=y, sr=sr, chunk_size=chunk_size, padding=padding, n_fft=n_fft, win_length=win_length,
            hop_length=hop_length, time_constant_s=time_constant_s, freq_mask_smooth_hz=freq_mask_smooth_hz,
            time_mask_smooth_ms=time_mask_smooth_ms, tmp_folder=tmp_folder, prop_decrease=prop_decrease,
            use_tqdm=use_tqdm, n_jobs=n_jobs,
        )
        self._lib = _lib                                   # tests: the CPU simulator build of the library
        if _lib is None:
            if not torch.cuda.is_available():
                raise RuntimeError("noisereduce_b200 has no CPU path: use_torch=True needs a CUDA device")
            self.device = torch.device(device)
            if self.device.type != "cuda":
                raise RuntimeError("noisereduce_b200 has no CPU path: device must be a CUDA device")
        else:
            self.device = torch.device("cpu")
        if y_noise is not None:                            # streamed_torch_gate.py:55-63
            y_noise = np.asarray(y_noise)
            if y_noise.shape[-1] > self.y.shape[-1] and clip_noise_stationary:
                y_noise = y_noise[..., : self.y.shape[-1]]
            y_noise = torch.from_numpy(np.ascontiguousarray(y_noise)).to(self.device)
            if y_noise.ndim == 1:
                y_noise = y_noise.unsqueeze(0)
        self.y_noise = y_noise
        self.tg = TorchGate(
            sr=sr,
            nonstationary=not stationary,
            n_std_thresh_stationary=n_std_thresh_stationary,
            n_thresh_nonstationary=thresh_n_mult_nonstationary,
            temp_coeff_nonstationary=1 / sigmoid_slope_nonstationary,
            n_movemean_nonstationary=int(time_constant_s / self._hop_length * sr),
            prop_decrease=prop_decrease,
            n_fft=self._n_fft,
            win_length=self._win_length,
            hop_length=self._hop_length,
            freq_mask_smooth_hz=freq_mask_smooth_hz,
            time_mask_smooth_ms=time_mask_smooth_ms,
        ).to(self.device)

    def _do_filter(self, chunk):
        """streamed_torch_gate.py:81-88: one padded chunk [C, Lp] through TorchGate; [C, (Lp // hop) * hop] back."""
        x = torch.from_numpy(np.ascontiguousarray(chunk)).to(self.device) if isinstance(chunk, np.ndarray) else chunk
        with torch.no_grad():
            out = self.tg(x=x, xn=self.y_noise, _lib=self._lib)
        return out.cpu().numpy()

    def get_traces(self, start_frame=None, end_frame=None):
        """base.py:167-226 as the reference runs it for this back-end: a loop over the chunk grid, each chunk padded,
        filtered by _do_filter and its centre copied out."""
        if start_frame is None:
            start_frame = 0
        if end_frame is None:
            end_frame = self.n_frames
        if self._chunk_size is not None and end_frame - start_frame > self._chunk_size:
            ich1, ich2 = int(start_frame / self._chunk_size), int((end_frame - 1) / self._chunk_size)
            out = np.zeros((self.n_channels, int(end_frame - start_frame)), dtype=self._dtype)
            pos = 0
            for ich in range(ich1, ich2 + 1):
                start0 = start_frame - ich * self._chunk_size if ich == ich1 else 0
                end0 = end_frame - ich * self._chunk_size if ich == ich2 else self._chunk_size
                part = self._get_filtered_chunk(ich)
                with np.errstate(invalid="ignore"):
                    out[:, pos: pos + end0 - start0] = part[:, start0:end0]
                pos += end0 - start0
        else:
            with np.errstate(invalid="ignore"):
                out = self.filter_chunk(start_frame=0, end_frame=end_frame).astype(self._dtype)
        return out.flatten() if self.flat else out
