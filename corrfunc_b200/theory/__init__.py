"""Drop-in counterparts of ``Corrfunc.theory.{DD,DDrppi,DDsmu,wp,xi,vpf}`` running on the GPU."""
from .pairs import DD, DDrppi, DDsmu, vpf, wp, xi

__all__ = ["DD", "DDrppi", "DDsmu", "wp", "xi", "vpf"]
