"""Corrfunc.theory.DDsmu -> corrfunc_b200.theory.DDsmu (GPU)."""
from corrfunc_b200.theory import DDsmu

__all__ = ["DDsmu"]
