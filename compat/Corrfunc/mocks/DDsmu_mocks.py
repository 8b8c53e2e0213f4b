"""Corrfunc.mocks.DDsmu_mocks -> corrfunc_b200.mocks.DDsmu_mocks (GPU)."""
from corrfunc_b200.mocks import DDsmu_mocks

__all__ = ["DDsmu_mocks"]
