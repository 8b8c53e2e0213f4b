"""Corrfunc.theory -> corrfunc_b200.theory (GPU).  As in the reference, the package attributes DD, DDrppi, ... are
the functions, and the sub-modules of the same names stay importable."""
from .DD import DD
from .DDrppi import DDrppi
from .DDsmu import DDsmu
from .vpf import vpf
from .wp import wp
from .xi import xi

__all__ = ["DD", "DDrppi", "DDsmu", "wp", "xi", "vpf"]
