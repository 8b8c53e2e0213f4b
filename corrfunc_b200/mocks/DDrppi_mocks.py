"""``Corrfunc.mocks.DDrppi_mocks`` look-alike (reference: Corrfunc/mocks/DDrppi_mocks.py:17-402)."""
from __future__ import annotations

import numpy as np

from .. import _capi, _lib
from ..utils import check_same_dtype, native_inputs, process_weights, translate_isa_string_to_enum
from .DDtheta_mocks import fix_ra_dec


def _mock_options(dtype, *, is_comoving_dist, verbose, need_avg, refine, max_cells_per_dim, copy_particles,
                  enable_min_sep_opt, c_api_timer, isa, fast_divide_and_NR_steps):
    translate_isa_string_to_enum(isa)
    custom = tuple(int(r) for r in refine) != (2, 2, 1)  # _countpairs_mocks.c:1200-1207
    opt = _capi.default_options(dtype, verbose=verbose, need_avg_sep=need_avg, bin_refine_factors=refine,
                                max_cells_per_dim=max_cells_per_dim, copy_particles=copy_particles,
                                enable_min_sep_opt=enable_min_sep_opt, c_api_timer=c_api_timer, isa=-1,
                                custom_refine=custom, is_comoving_dist=is_comoving_dist)
    opt.fast_divide_and_NR_steps = int(fast_divide_and_NR_steps)
    return opt


def DDrppi_mocks(autocorr, cosmology, nthreads, pimax, binfile, RA1, DEC1, CZ1, weights1=None, RA2=None, DEC2=None,
                 CZ2=None, weights2=None, is_comoving_dist=False, verbose=False, output_rpavg=False,
                 fast_divide_and_NR_steps=0, xbin_refine_factor=2, ybin_refine_factor=2, zbin_refine_factor=1,
                 max_cells_per_dim=100, copy_particles=True, enable_min_sep_opt=True, c_api_timer=False,
                 isa="fastest", weight_type=None):
    """Survey-geometry pair counts DD(rp, pi) from RA, DEC (degrees) and CZ (km/s; or the comoving distance with
    ``is_comoving_dist=True``), line of sight = pair midpoint.  ``cosmology``: 1 (LasDamas) or 2 (Planck).  Returns a structured array (rmin, rmax, rpavg, pimax, npairs, weightavg), rp-major with
    ``int(pimax)`` unit-width pi bins [and the C call's wall time when ``c_api_timer``]."""
    if not autocorr and (RA2 is None or DEC2 is None or CZ2 is None):
        raise ValueError("Must pass valid arrays for RA2/DEC2/CZ2 for computing cross-correlation")
    (RA1, DEC1, CZ1, RA2, DEC2, CZ2), weights1, weights2, dtype = native_inputs((RA1, DEC1, CZ1, RA2, DEC2, CZ2), weights1, weights2, RA1, RA2, weight_type, autocorr)
    RA1, DEC1 = fix_ra_dec(RA1, DEC1)
    if autocorr == 0:
        RA2, DEC2 = fix_ra_dec(RA2, DEC2)
    opt = _mock_options(dtype, is_comoving_dist=is_comoving_dist, verbose=verbose, need_avg=output_rpavg,
                        refine=(xbin_refine_factor, ybin_refine_factor, zbin_refine_factor),
                        max_cells_per_dim=max_cells_per_dim, copy_particles=copy_particles,
                        enable_min_sep_opt=enable_min_sep_opt, c_api_timer=c_api_timer, isa=isa,
                        fast_divide_and_NR_steps=fast_divide_and_NR_steps)
    w1 = None if weights1 is None else np.ascontiguousarray(weights1[0])
    w2 = None if weights2 is None else np.ascontiguousarray(weights2[0])
    r = _capi.call_DDrppi_mocks(_lib.load(), autocorr, cosmology, nthreads, pimax, binfile, RA1, DEC1, CZ1, w1=w1,
                                RA2=RA2, DEC2=DEC2, CZ2=CZ2, w2=w2, weight_type=weight_type, options=opt, dtype=dtype)
    nrp, npi = r["npairs"].shape
    if r["npairs"].size == 0:  # empty particle set: an empty table, like the reference
        npi = max(npi, 1)
    res = np.zeros(nrp * npi, dtype=[("rmin", np.float64), ("rmax", np.float64), ("rpavg", np.float64),
                                     ("pimax", np.float64), ("npairs", np.uint64), ("weightavg", np.float64)])
    dpi = r["pimax"] / npi  # rows as built in _countpairs_mocks.c (rp-major, upper pi edge per row)
    res["rmin"] = np.repeat(r["rupp"][:-1], npi)
    res["rmax"] = np.repeat(r["rupp"][1:], npi)
    res["pimax"] = np.tile((np.arange(npi) + 1) * dpi, nrp)
    res["rpavg"], res["npairs"], res["weightavg"] = r["ravg"].ravel(), r["npairs"].ravel(), r["weightavg"].ravel()
    return (res, r["api_time"]) if c_api_timer else res
