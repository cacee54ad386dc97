"""The edge cases of reduce_noise() pinned by tests/golden/edge_cases.npz (shared by the generator and the tests)."""
import numpy as np

from tests.synth_host import synth_small

SR = 16000


def cases():
    """name -> (y, kwargs).  Inputs are rebuilt from the seeded generator, only the reference's OUTPUTS are stored."""
    base = synth_small(C=2, n=30000)
    c = {}
    c["short_nonstationary_700"] = (base[0, :700], dict(stationary=False))
    c["len_eq_n_fft"] = (base[0, :1024], dict(stationary=True))
    # stationary.py:67-73 through scipy.signal.stft: a noise clip shorter than the window shrinks the window to the clip
    c["clip_900_shorter_than_window"] = (base[0, :900], dict(stationary=True))
    c["clip_800_two_channels"] = (base[:, :800], dict(stationary=True))
    c["y_noise_900"] = (base[:, :8000], dict(stationary=True, y_noise=base[0, :900]))
    c["chunk_size_900_clips_the_noise"] = (base[:, :8000], dict(stationary=True, chunk_size=900, padding=100))
    c["chunk_larger_than_signal"] = (base[:, :9000], dict(stationary=True, chunk_size=20000, padding=100))
    c["padding_zero"] = (base[:, :9000], dict(stationary=True, chunk_size=3000, padding=0))
    c["ragged_last_chunk_nonstationary"] = (base[:, :10001], dict(stationary=False, chunk_size=3333, padding=777))
    c["fortran_order"] = (np.asfortranarray(base[:, :8000]), dict(stationary=True, chunk_size=4000, padding=300))
    c["strided_view"] = (base[:, :16000:2], dict(stationary=True, chunk_size=4000, padding=300))
    c["float16"] = (base[:, :8000].astype(np.float16), dict(stationary=True, chunk_size=4000, padding=300))
    c["int32"] = ((base[:, :8000] * 20000).astype(np.int32), dict(stationary=True, chunk_size=4000, padding=300))
    c["uint8"] = ((base[:, :8000] * 100 + 128).astype(np.uint8), dict(stationary=True, chunk_size=4000, padding=300))
    c["zeros_nonstationary_is_nan"] = (np.zeros((1, 6000), np.float32), dict(stationary=False))
    c["constant"] = (np.full((1, 6000), 0.25, np.float32), dict(stationary=True))
    c["n_std_zero"] = (base[:, :8000], dict(stationary=True, n_std_thresh_stationary=0.0, chunk_size=4000, padding=300))
    c["no_clip_noise"] = (base[:, :12000], dict(stationary=True, clip_noise_stationary=False, chunk_size=4000, padding=300))
    return c


# inputs the reference refuses, with the exception type (the message is matched where the reference's own is meaningful)
def refused():
    base = synth_small(C=2, n=30000)
    return {
        "clip_768_not_longer_than_noverlap": (base[0, :768], dict(stationary=True), ValueError, "noverlap must be less than nperseg"),
        "y_noise_500": (base[:, :8000], dict(stationary=True, y_noise=base[0, :500]), ValueError, "noverlap must be less than nperseg"),
        "complex_samples": (base[:, :8000].astype(np.complex64), dict(stationary=True), ValueError, None),
        "empty": (np.zeros((1, 0), np.float32), dict(stationary=True), ValueError, "zero-size array"),
    }
