"""Drop-in counterpart of ``Corrfunc.mocks.DDtheta_mocks`` running on the GPU."""
from .DDtheta_mocks import DDtheta_mocks

__all__ = ["DDtheta_mocks"]
