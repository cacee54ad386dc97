"""TorchGate surface on the CPU simulator build (kernel logic + host mirror), vs oracle/torchgate_oracle."""
import numpy as np
import pytest
import torch

from noisereduce_b200.torchgate import TorchGate
from oracle import torchgate_oracle as TO
from tests.cusim_util import cusim_library
from tests.synth_host import synth_torchgate

TOL = 5e-6     # FP32 pipeline vs the float64 oracle (no mask decision sits at its threshold in these inputs)


def rel(a, b):
    return float(np.abs(a - b).max() / np.abs(b).max())


@pytest.fixture(scope="module")
def lib():
    return cusim_library()


@pytest.mark.parametrize("kw", [
    dict(), dict(prop_decrease=0.7), dict(nonstationary=True),
    dict(nonstationary=True, prop_decrease=0.6, n_movemean_nonstationary=7, n_thresh_nonstationary=1.0),
    dict(freq_mask_smooth_hz=None, time_mask_smooth_ms=None),
])
def test_forward_matches_oracle(lib, kw):
    x = synth_torchgate(B=2, n=6000)
    w = torch.hann_window(1024).numpy()
    tg = TorchGate(sr=16000, **kw)
    y = tg(torch.from_numpy(x), _lib=lib)
    assert y.dtype == torch.float32 and tuple(y.shape) == (2, (6000 // 256) * 256)
    ref = TO.torchgate_forward(x.astype(np.float64), 16000, window=w, **kw)
    assert rel(y.numpy(), ref) < TOL


def test_xn_and_dtypes_and_module_contract(lib):
    x = synth_torchgate(B=2, n=6000)
    w = torch.hann_window(1024).numpy()
    tg = TorchGate(sr=16000)
    for xn in (x[:1, :3000], x[0, :3000], x[:, 1000:4000]):               # [1, Ln], [Ln], [B, Ln]
        y = tg(torch.from_numpy(x), torch.from_numpy(xn), _lib=lib).numpy()
        ref = TO.torchgate_forward(x.astype(np.float64), 16000, xn=xn.astype(np.float64), window=w)
        assert rel(y, ref) < TOL
    y64 = tg(torch.from_numpy(x).double(), _lib=lib)
    assert y64.dtype == torch.float64
    assert list(tg.state_dict().keys()) == ["smoothing_filter"] and tuple(tg.smoothing_filter.shape) == (1, 1, 33, 7)
    assert TorchGate(sr=16000, freq_mask_smooth_hz=None, time_mask_smooth_ms=None).smoothing_filter is None
    with pytest.raises(Exception, match="x must be bigger than 2048"):
        tg(torch.zeros(1, 1000), _lib=lib)
    with pytest.raises(AssertionError):
        tg(torch.zeros(1000), _lib=lib)
    with pytest.raises(AssertionError):
        TorchGate(sr=16000, prop_decrease=1.5)
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        tg(torch.zeros(1, 4096))                                           # product path: no CPU fallback
    with pytest.raises(NotImplementedError, match="no backward pass"):
        tg(torch.from_numpy(x).requires_grad_(True), _lib=lib)             # never a silently detached result
    with torch.no_grad():
        assert tg(torch.from_numpy(x).requires_grad_(True), _lib=lib).shape[0] == 2


@pytest.mark.parametrize("geo", [
    dict(n_fft=512), dict(n_fft=512, win_length=400, hop_length=100), dict(n_fft=256, win_length=255, hop_length=50),
    dict(n_fft=2048), dict(n_fft=400), dict(n_fft=441),
], ids=lambda g: "-".join(f"{k}{v}" for k, v in g.items()))
def test_non_default_geometry_runs_on_the_general_family(lib, geo):
    """TorchGate off n_fft=1024/hop=256: float64 general family with torch.stft framing (gate_generic.cuh)."""
    x = synth_torchgate(B=2, n=6000)
    w = torch.hann_window(geo.get("win_length", geo["n_fft"])).numpy()
    for kw in (dict(prop_decrease=0.7), dict(nonstationary=True, n_movemean_nonstationary=7),
               dict(freq_mask_smooth_hz=None, time_mask_smooth_ms=None)):
        tg = TorchGate(sr=16000, **geo, **kw)
        y = tg(torch.from_numpy(x), _lib=lib)
        ref = TO.torchgate_forward(x.astype(np.float64), 16000, window=w, **geo, **kw)
        assert tuple(y.shape) == ref.shape and rel(y.numpy(), ref) < 5e-7
    tg = TorchGate(sr=16000, **geo)
    for xn in (x[0, :5000], x[:, 500:5500]):
        y = tg(torch.from_numpy(x), torch.from_numpy(xn), _lib=lib).numpy()
        assert rel(y, TO.torchgate_forward(x.astype(np.float64), 16000, xn=xn.astype(np.float64), window=w, **geo)) < 5e-7
