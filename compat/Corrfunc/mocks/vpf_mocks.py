"""Corrfunc.mocks.vpf_mocks -> corrfunc_b200.mocks.vpf_mocks (GPU)."""
from corrfunc_b200.mocks import vpf_mocks

__all__ = ["vpf_mocks"]
