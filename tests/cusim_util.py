"""TEST INFRASTRUCTURE: load the CPU-simulator build of libb200gate (tests/cusim)."""
import functools

from noisereduce_b200 import _cabi


@functools.lru_cache(maxsize=1)
def cusim_library():
    from tests.cusim import build_cusim
    return _cabi.GateLibrary(build_cusim.build())
