"""Corrfunc.mocks -> corrfunc_b200.mocks (GPU)."""
from .DDtheta_mocks import DDtheta_mocks

__all__ = ["DDtheta_mocks"]
