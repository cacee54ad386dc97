"""Mirror of noisereduce/spectralgate/stationary.py:7 (SpectralGateStationary) on libb200gate."""
import numpy as np

from .. import _cabi
from .base import SpectralGate


class SpectralGateStationary(SpectralGate):
    def __init__(
        self,
        y,
        sr,
        y_noise,
        n_std_thresh_stationary,
        chunk_size,
        clip_noise_stationary,
        padding,
        n_fft,
        win_length,
        hop_length,
        time_constant_s,
        freq_mask_smooth_hz,
        time_mask_smooth_ms,
        tmp_folder,
        prop_decrease,
        use_tqdm,
        n_jobs,
    ):
        super().__init__(
            y=y, sr=sr, chunk_size=chunk_size, padding=padding, n_fft=n_fft, win_length=win_length,
            hop_length=hop_length, time_constant_s=time_constant_s, freq_mask_smooth_hz=freq_mask_smooth_hz,
            time_mask_smooth_ms=time_mask_smooth_ms, tmp_folder=tmp_folder, prop_decrease=prop_decrease,
            use_tqdm=use_tqdm, n_jobs=n_jobs,
        )
        self.n_std_thresh_stationary = n_std_thresh_stationary

        if y_noise is None:                                     # stationary.py:47-59
            noise = self.y
        else:
            y_noise = np.array(y_noise)
            if len(y_noise.shape) == 1:
                noise = np.expand_dims(y_noise, 0)
            elif len(y.shape) > 2:
                raise ValueError("Waveform must be in shape (# frames, # channels)")
            else:
                noise = y_noise

        self._noise2d, self._clip_noise = noise, bool(clip_noise_stationary)
        params = self._gate_params()
        params.update(stationary=1, n_std_thresh=float(n_std_thresh_stationary),
                      clip_noise=1 if clip_noise_stationary else 0)
        self._params = params
        self._gate = _cabi.Gate(**params)
        # stationary.py:61-81 on the device: channel mean in the input dtype, clip, STFT, dB with the
        # 80 dB floor, per-bin mean / std over time, threshold
        clip = noise
        if noise.dtype not in (np.float32, np.int16, np.float64):
            # the reference collapses the clip with np.mean in the clip's own dtype rules (stationary.py:61): float16 stays
            # float16 (float32 accumulation, rounded), integers / bool become float64 -- numpy decides, the device gets the result
            clip = np.mean(noise, axis=0)[None, :]
            if clip.dtype not in (np.float32, np.float64):
                clip = clip.astype(np.float32)
        n_eff = clip.shape[1]
        if clip_noise_stationary and self._chunk_size is not None:
            n_eff = min(n_eff, int(self._chunk_size))                         # stationary.py:63-64
        noverlap = self._win_length - self._hop_length
        if n_eff < self._win_length:
            # scipy.signal.stft on a clip shorter than the window shrinks the window to the clip (nperseg = len(x), with a
            # warning) and then insists on noverlap < nperseg; the reference inherits both behaviours (stationary.py:67-73)
            if noverlap >= n_eff:
                raise ValueError("noverlap must be less than nperseg.")
            short = dict(params)
            short.update(win_length=int(n_eff), hop_length=int(n_eff - noverlap))
            stats_gate = _cabi.Gate(**short)
            try:
                stats_gate.noise_stats_host(self._samples_for_device(clip[:, :n_eff]))
                self.mean_freq_noise, self.std_freq_noise = stats_gate.noise_mean_std()
                self.noise_thresh = stats_gate.noise_threshold()
            finally:
                stats_gate.close()
            self._gate.set_noise_threshold(self.noise_thresh)
        else:
            self._gate.noise_stats_host(self._samples_for_device(clip))
            self.mean_freq_noise, self.std_freq_noise = self._gate.noise_mean_std()
            self.noise_thresh = self._gate.noise_threshold()

    def _unit_gate_params(self):
        p = dict(self._params)
        p.update(chunk_size=0, padding=0)
        return p

    def _prepare_unit_gate(self, gate):
        gate.set_noise_threshold(self.noise_thresh)          # the thresholds of the whole recording (stationary.py:79-81)

    @property
    def y_noise(self):
        """The collapsed noise clip the reference keeps as an attribute (stationary.py:61-64): channel mean in
        the input dtype, clipped to chunk_size.  Computed on demand -- the thresholds come from the device."""
        n = self._noise2d[:, : self._chunk_size] if self._clip_noise else self._noise2d
        return np.mean(n, axis=0)
