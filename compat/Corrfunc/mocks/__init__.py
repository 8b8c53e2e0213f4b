"""Corrfunc.mocks -> corrfunc_b200.mocks (GPU)."""
from .DDrppi_mocks import DDrppi_mocks
from .DDsmu_mocks import DDsmu_mocks
from .DDtheta_mocks import DDtheta_mocks
from .vpf_mocks import vpf_mocks

__all__ = ["DDtheta_mocks", "DDrppi_mocks", "DDsmu_mocks", "vpf_mocks"]
