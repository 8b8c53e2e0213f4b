"""Corrfunc.mocks.DDtheta_mocks -> corrfunc_b200.mocks.DDtheta_mocks (GPU)."""
from corrfunc_b200.mocks import DDtheta_mocks

__all__ = ["DDtheta_mocks"]
