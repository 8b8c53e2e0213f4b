"""Corrfunc.theory.vpf -> corrfunc_b200.theory.vpf (GPU)."""
from corrfunc_b200.theory import vpf

__all__ = ["vpf"]
