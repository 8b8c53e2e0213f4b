"""Corrfunc.io -> corrfunc_b200.io."""
from corrfunc_b200.io import read_ascii_catalog, read_catalog, read_fastfood_catalog

__all__ = ["read_fastfood_catalog", "read_ascii_catalog", "read_catalog"]
