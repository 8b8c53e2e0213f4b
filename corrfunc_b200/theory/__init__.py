"""Drop-in counterparts of ``Corrfunc.theory.{DD,DDrppi,DDsmu,wp,xi}`` running on the GPU."""
from .pairs import DD, DDrppi, DDsmu, wp, xi

__all__ = ["DD", "DDrppi", "DDsmu", "wp", "xi"]
