"""Import-path shim: put this directory's parent (``compat/``) on PYTHONPATH and scripts written against
manodeep/Corrfunc (``from Corrfunc.theory.DD import DD`` ...) run on corrfunc_b200 unchanged:
theory.{DD,DDrppi,DDsmu,wp,xi,vpf}, mocks.{DDtheta_mocks,DDrppi_mocks,DDsmu_mocks,vpf_mocks}, utils, io and the
three small helpers of the reference's top-level package."""
import os
import shutil
import sys

_root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _root not in sys.path:
    sys.path.insert(0, _root)

__version__ = "2.5.3"  # the API version this shim mirrors
__all__ = ["theory", "mocks", "utils", "io", "read_text_file", "write_text_file", "which"]


def read_text_file(filename, encoding="utf-8"):
    """The contents of a text file as one string (Corrfunc/__init__.py:36-49)."""
    with open(filename, "r", encoding=encoding) as f:
        return f.read()


def write_text_file(filename, contents, encoding="utf-8"):
    """Write ``contents`` to a text file (Corrfunc/__init__.py:52-65)."""
    with open(filename, "w", encoding=encoding) as f:
        f.write(contents)


def which(program, mode=os.F_OK | os.X_OK, path=None):
    """Full path of an executable, or None (Corrfunc/__init__.py:68-110; ``shutil.which``)."""
    return shutil.which(program, mode, path)
