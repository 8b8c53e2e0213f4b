"""Corrfunc.theory.DDrppi -> corrfunc_b200.theory.DDrppi (GPU)."""
from corrfunc_b200.theory import DDrppi

__all__ = ["DDrppi"]
