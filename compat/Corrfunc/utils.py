"""Corrfunc.utils: the post-processing estimators (the rest of the reference module is not mirrored)."""
from corrfunc_b200.utils import convert_3d_counts_to_cf, convert_rp_pi_counts_to_wp

__all__ = ["convert_3d_counts_to_cf", "convert_rp_pi_counts_to_wp"]
