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
    with torch.no_grad():
        assert tg(torch.from_numpy(x).requires_grad_(True), _lib=lib).shape[0] == 2
    ref64 = TO.torchgate_forward(x.astype(np.float64), 16000, window=w)
    assert rel(y64.numpy(), ref64) < 5e-7            # float64 input runs the float64 general family (complex128 like the reference)


@pytest.mark.parametrize("kw", [dict(), dict(nonstationary=True, prop_decrease=0.8)])
def test_backward_is_the_adjoint_with_the_forward_masks(lib, kw):
    """torchgate.py:223-262: masks under no_grad, gradient through stft -> * mask -> istft.  The CUDA path's backward is
    env * gate(g / env) with the stored masks (b200gate_torch_apply_masks).  Checked without the reference: (1) the
    forward is unchanged by requires_grad, (2) <A v, g> == <v, A^T g> for the linear map A = gate with these masks,
    (3) A x == forward(x)."""
    B, L = 2, 24 * 256
    x = torch.from_numpy(synth_torchgate(B=B, n=L)).requires_grad_(True)
    tg = TorchGate(sr=16000, **kw)
    y = tg(x, _lib=lib)
    assert y.requires_grad and tuple(y.shape) == (B, L)
    with torch.no_grad():
        y0 = tg(x.detach(), _lib=lib)
    assert torch.equal(y.detach(), y0)
    tg2 = TorchGate(sr=16000, **kw)                    # fresh module so that its masks are those of x
    y = tg2(x, _lib=lib)
    gen = torch.Generator().manual_seed(3)
    g = torch.randn(y.shape, generator=gen)
    (y * g).sum().backward()
    grad = x.grad.clone()
    gate = next(iter(tg2._gate.values()))
    v = torch.randn((B, L), generator=gen) * 0.05
    Av = torch.empty((B, L))
    gate.torch_apply_masks_device(v.data_ptr(), Av.data_ptr(), np.float32, B, L, L, L, None)
    lhs, rhs = float((Av.double() * g.double()).sum()), float((v.double() * grad.double()).sum())
    assert abs(lhs - rhs) <= 2e-5 * max(abs(lhs), abs(rhs), 1e-3)
    Ax = torch.empty((B, L))
    gate.torch_apply_masks_device(x.detach().contiguous().data_ptr(), Ax.data_ptr(), np.float32, B, L, L, L, None)
    assert rel(Ax.numpy(), y.detach().numpy()) < 1e-6
    # a second forward overwrites the masks: backward of the first must refuse
    ya = tg2(x, _lib=lib)
    tg2(x.detach() * 0.5, _lib=lib)
    with pytest.raises(RuntimeError, match="another forward ran"):
        ya.sum().backward()
    with pytest.raises(NotImplementedError, match="backward pass exists"):
        TorchGate(sr=16000, n_fft=512)(x, _lib=lib)


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


def test_streamed_torch_gate_route_of_reduce_noise(lib):
    """reduce_noise(use_torch=True) (noisereduce.py:121-143 -> StreamedTorchGate): every padded chunk through TorchGate
    with the reference's parameter mapping (streamed_torch_gate.py:66-79), assembled by the chunk loop of base.py:167-226.
    Expected values: the torch-surface oracle applied chunk by chunk."""
    from noisereduce_b200.spectralgate.streamed_torch_gate import StreamedTorchGate
    from oracle import spectral_gate_oracle as O
    y = synth_torchgate(B=2, n=9000)
    w = torch.hann_window(1024).numpy()
    sr, cs, pad = 16000, 3000, 1100
    for stationary in (True, False):
        sg = StreamedTorchGate(y=y, sr=sr, stationary=stationary, chunk_size=cs, padding=pad, time_constant_s=0.2,
                               thresh_n_mult_nonstationary=1.5, sigmoid_slope_nonstationary=8, _lib=lib)
        got = sg.get_traces()
        assert got.shape == y.shape and got.dtype == y.dtype
        want = np.zeros(y.shape)
        kw = dict(nonstationary=not stationary, n_thresh_nonstationary=1.5, temp_coeff_nonstationary=1 / 8,
                  n_movemean_nonstationary=int(0.2 / 256 * sr))
        for (i1, i2, lo, hi) in O.chunk_table(y.shape[1], cs, pad):
            part = TO.torchgate_forward(O.read_chunk(y, i1, i2), sr, window=w, **kw)
            want[:, lo:hi] = part[:, lo - i1: hi - i1]
        assert rel(got, want) < TOL
    # a recording shorter than chunk_size is one padded chunk (base.py:222); flat input stays flat
    sg = StreamedTorchGate(y=y[0, :2500], sr=sr, stationary=True, chunk_size=cs, padding=pad, _lib=lib)
    got = sg.get_traces()
    part = TO.torchgate_forward(O.read_chunk(y[:1, :2500], -pad, 2500 + pad), sr, window=w)
    assert got.shape == (2500,) and rel(got, part[0, pad: pad + 2500]) < TOL


@pytest.mark.parametrize("name,kw", [("stat", {}), ("nonstat", dict(nonstationary=True, prop_decrease=0.8))])
def test_backward_matches_the_reference_gradient(lib, name, kw):
    """tests/golden/torchgate_grad.npz (make_golden_grad.py: the unmodified reference's autograd on CPU).  The input is
    100 samples longer than a multiple of hop: the gradient on the output's span is compared; the samples beyond the
    output get zero here (documented in TorchGate._adjoint) where the reference leaks a small boundary term."""
    import os
    gd = np.load(os.path.join(os.path.dirname(__file__), "golden", "torchgate_grad.npz"))
    x = torch.from_numpy(gd["x"]).requires_grad_(True)
    tg = TorchGate(sr=int(gd["sr"]), **kw)
    y = tg(x, _lib=lib)
    assert rel(y.detach().numpy(), gd[f"{name}_y"]) < 2e-5
    (y * torch.from_numpy(gd[f"{name}_g"])).sum().backward()
    Lo = y.shape[1]
    want = gd[f"{name}_grad"]
    assert float(np.abs(x.grad.numpy()[:, :Lo] - want[:, :Lo]).max()) < 2e-5 * float(np.abs(want).max())
    assert np.all(x.grad.numpy()[:, Lo:] == 0)
