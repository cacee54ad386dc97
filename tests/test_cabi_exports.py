"""CPU-side checks of the product library: it builds for sm_100a, loads, exports every symbol the
header declares, and fails loudly (no fallback) when there is no CUDA device."""
import ctypes
import os
import re

import pytest

from noisereduce_b200 import _cabi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def product_lib():
    from noisereduce_b200.csrc import build
    return build.build()


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "b200gate.h")).read()
    return sorted(set(re.findall(r"\b(b200gate_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree():
    assert declared_symbols() == sorted(_cabi.EXPORTS)


def test_library_exports_every_declared_symbol(product_lib):
    dll = ctypes.CDLL(product_lib)
    for sym in declared_symbols():
        assert hasattr(dll, sym), sym


def test_params_struct_layout_matches_header():
    # field order and the 8-byte members' alignment are what the C side reads
    names = [f for f, _ in _cabi.Params._fields_]
    src = open(os.path.join(ROOT, "include", "b200gate.h")).read()
    body = src[src.index("typedef struct b200gate_params {"): src.index("} b200gate_params;")]
    c_names = re.findall(r"^\s*(?:int32_t|int64_t|double)\s+([a-z0-9_]+);", body, flags=re.M)
    assert names == c_names
    assert ctypes.sizeof(_cabi.Params) == 14 * 4 + 2 * 8 + 8 * 8
    names = [f for f, _ in _cabi.Stats._fields_]
    body = src[src.index("typedef struct b200gate_stats {"): src.index("} b200gate_stats;")]
    c_names = []
    for line in re.findall(r"^\s*(?:int64_t|double)\s+([^;]+);", body, flags=re.M):
        c_names += [n.strip() for n in line.split(",")]
    assert names == c_names


def test_no_cpu_fallback_without_gpu(product_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the no-device error path cannot be exercised")
    lib = _cabi.GateLibrary(product_lib)
    with pytest.raises(_cabi.GateError, match="no CUDA device|CUDA"):
        _cabi.Gate(lib=lib, surface=0, stationary=1, n_fft=1024, win_length=1024, hop_length=256,
                   n_grad_freq=5, n_grad_time=9, chunk_size=600000, padding=30000, sr=48000.0,
                   prop_decrease=1.0, n_std_thresh=1.5, top_db=80.0)


def test_missing_library_is_a_loud_error(tmp_path):
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _cabi.GateLibrary(str(tmp_path / "nope.so"))
