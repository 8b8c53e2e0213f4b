"""Import-path shim: put this directory's parent (``compat/``) on PYTHONPATH and scripts written against
manodeep/Corrfunc (``from Corrfunc.theory.DD import DD`` ...) run on corrfunc_b200 unchanged, for the hot path this
repository covers: theory.{DD,DDrppi,DDsmu,wp,xi}, mocks.DDtheta_mocks and the two estimators of Corrfunc.utils.
Anything else the reference offers (vpf, DDrppi_mocks, DDsmu_mocks, io) is not provided and raises ImportError."""
import os
import sys

_root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _root not in sys.path:
    sys.path.insert(0, _root)

__version__ = "2.5.3"  # the API version this shim mirrors
__all__ = ["theory", "mocks", "utils"]
