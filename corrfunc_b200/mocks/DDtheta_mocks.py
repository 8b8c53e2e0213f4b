"""``Corrfunc.mocks.DDtheta_mocks`` look-alike (reference: Corrfunc/mocks/DDtheta_mocks.py:19-360)."""
from __future__ import annotations

import numpy as np

from .. import _capi, _lib
from ..utils import check_same_dtype, native_inputs, process_weights, translate_isa_string_to_enum


def fix_ra_dec(ra, dec):
    """RA -> [0,360], DEC -> [-90,90] (Corrfunc/utils.py:421-461); returns copies."""
    ra = np.array(ra, copy=True)
    dec = np.array(dec, copy=True)
    if ra.size and ra.min() < 0.0:
        print("Warning: found negative RA values, wrapping into [0.0, 360.0]  range")
        ra += 180.0
    if dec.size and dec.max() > 90.0:
        print("Warning: found DEC values more than 90.0; wrapping into [-90.0, 90.0] range")
        dec -= 90.0
    return ra, dec


def DDtheta_mocks(autocorr, nthreads, binfile, RA1, DEC1, weights1=None, RA2=None, DEC2=None, weights2=None,
                  link_in_dec=True, link_in_ra=True, verbose=False, output_thetaavg=False, fast_acos=False,
                  ra_refine_factor=2, dec_refine_factor=2, max_cells_per_dim=100, copy_particles=True,
                  enable_min_sep_opt=True, c_api_timer=False, isa="fastest", weight_type=None):
    """Angular pair counts DD(theta) for points on the sky (degrees).  Returns a structured array
    (thetamin, thetamax, thetaavg, npairs, weightavg) [and the C call's wall time when
    ``c_api_timer``]."""
    if autocorr == 0 and (RA2 is None or DEC2 is None):
        raise ValueError("Must pass valid arrays for RA2/DEC2 for computing cross-correlation")
    if link_in_ra and not link_in_dec:
        raise ValueError("Linking in RA requires linking in DEC as well")  # mocks.options: LINK_IN_RA needs LINK_IN_DEC
    translate_isa_string_to_enum(isa)
    (RA1, DEC1, RA2, DEC2), weights1, weights2, dtype = native_inputs((RA1, DEC1, RA2, DEC2), weights1, weights2, RA1, RA2, weight_type, autocorr)
    RA1, DEC1 = fix_ra_dec(RA1, DEC1)
    if autocorr == 0:
        RA2, DEC2 = fix_ra_dec(RA2, DEC2)
    refine = (ra_refine_factor, dec_refine_factor, 1)
    custom = (int(ra_refine_factor), int(dec_refine_factor)) != (2, 2)
    opt = _capi.default_options(dtype, verbose=verbose, need_avg_sep=output_thetaavg, bin_refine_factors=refine,
                                max_cells_per_dim=max_cells_per_dim, copy_particles=copy_particles,
                                enable_min_sep_opt=enable_min_sep_opt, c_api_timer=c_api_timer, isa=-1,
                                link_in_dec=link_in_dec, link_in_ra=link_in_ra, fast_acos=fast_acos,
                                custom_refine=custom)
    w1 = None if weights1 is None else np.ascontiguousarray(weights1[0])
    w2 = None if weights2 is None else np.ascontiguousarray(weights2[0])
    r = _capi.call_DDtheta(_lib.load(), autocorr, nthreads, binfile, RA1, DEC1, w1=w1, RA2=RA2, DEC2=DEC2, w2=w2,
                           weight_type=weight_type, options=opt, dtype=dtype)
    res = np.zeros(r["npairs"].size, dtype=[("thetamin", np.float64), ("thetamax", np.float64),
                                            ("thetaavg", np.float64), ("npairs", np.uint64),
                                            ("weightavg", np.float64)])
    res["thetamin"], res["thetamax"] = r["rupp"][:-1], r["rupp"][1:]
    res["thetaavg"], res["npairs"], res["weightavg"] = r["ravg"], r["npairs"], r["weightavg"]
    return (res, r["api_time"]) if c_api_timer else res
