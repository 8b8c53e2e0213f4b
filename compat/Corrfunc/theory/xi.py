"""Corrfunc.theory.xi -> corrfunc_b200.theory.xi (GPU)."""
from corrfunc_b200.theory import xi

__all__ = ["xi"]
