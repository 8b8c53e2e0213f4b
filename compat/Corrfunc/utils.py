"""Corrfunc.utils -> corrfunc_b200.utils."""
from corrfunc_b200.utils import (compute_nbins, convert_3d_counts_to_cf, convert_rp_pi_counts_to_wp,
                                 convert_to_native_endian, fix_cz, fix_ra_dec, gridlink_sphere, is_native_endian,
                                 process_weights, return_file_with_rbins, sys_pipes, translate_isa_string_to_enum)

__all__ = ["convert_3d_counts_to_cf", "convert_rp_pi_counts_to_wp", "return_file_with_rbins", "fix_cz", "fix_ra_dec",
           "translate_isa_string_to_enum", "compute_nbins", "gridlink_sphere", "convert_to_native_endian",
           "is_native_endian", "process_weights", "sys_pipes"]
