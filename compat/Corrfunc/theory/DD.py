"""Corrfunc.theory.DD -> corrfunc_b200.theory.DD (GPU)."""
from corrfunc_b200.theory import DD

__all__ = ["DD"]
