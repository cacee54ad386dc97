from .nonstationary import SpectralGateNonStationary
from .stationary import SpectralGateStationary

__all__ = ["SpectralGateStationary", "SpectralGateNonStationary"]
