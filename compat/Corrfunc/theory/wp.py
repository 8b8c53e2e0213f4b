"""Corrfunc.theory.wp -> corrfunc_b200.theory.wp (GPU)."""
from corrfunc_b200.theory import wp

__all__ = ["wp"]
