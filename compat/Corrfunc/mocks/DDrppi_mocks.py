"""Corrfunc.mocks.DDrppi_mocks -> corrfunc_b200.mocks.DDrppi_mocks (GPU)."""
from corrfunc_b200.mocks import DDrppi_mocks

__all__ = ["DDrppi_mocks"]
