"""``Corrfunc.mocks.DDsmu_mocks`` look-alike (reference: Corrfunc/mocks/DDsmu_mocks.py:17-385)."""
from __future__ import annotations

import numpy as np

from .. import _capi, _lib
from ..utils import check_same_dtype, native_inputs, process_weights
from .DDrppi_mocks import _mock_options
from .DDtheta_mocks import fix_ra_dec


def DDsmu_mocks(autocorr, cosmology, nthreads, mu_max, nmu_bins, binfile, RA1, DEC1, CZ1, weights1=None, RA2=None,
                DEC2=None, CZ2=None, weights2=None, is_comoving_dist=False, verbose=False, output_savg=False,
                fast_divide_and_NR_steps=0, xbin_refine_factor=2, ybin_refine_factor=2, zbin_refine_factor=1,
                max_cells_per_dim=100, copy_particles=True, enable_min_sep_opt=True, c_api_timer=False,
                isa="fastest", weight_type=None):
    """Survey-geometry pair counts DD(s, mu) from RA, DEC (degrees) and CZ (km/s; or the comoving distance with
    ``is_comoving_dist=True``), mu measured against the pair-midpoint line of sight.  Returns a structured array (smin, smax, savg, mumax, npairs, weightavg), s-major
    with ``nmu_bins`` mu bins up to ``mu_max`` [and the C call's wall time when ``c_api_timer``]."""
    if not autocorr and (RA2 is None or DEC2 is None or CZ2 is None):
        raise ValueError("Must pass valid arrays for RA2/DEC2/CZ2 for computing cross-correlation")
    if mu_max <= 0.0 or mu_max > 1.0:  # DDsmu_mocks.py: "The parameter `mu_max` (= ...) must be > 0 and <= 1.0"
        raise ValueError("The parameter `mu_max` (= {0}), the max. of cosine of the angle to the line-of-sight (LOS), "
                         "must be > 0 and <= 1.0".format(mu_max))
    (RA1, DEC1, CZ1, RA2, DEC2, CZ2), weights1, weights2, dtype = native_inputs((RA1, DEC1, CZ1, RA2, DEC2, CZ2), weights1, weights2, RA1, RA2, weight_type, autocorr)
    RA1, DEC1 = fix_ra_dec(RA1, DEC1)
    if autocorr == 0:
        RA2, DEC2 = fix_ra_dec(RA2, DEC2)
    opt = _mock_options(dtype, is_comoving_dist=is_comoving_dist, verbose=verbose, need_avg=output_savg,
                        refine=(xbin_refine_factor, ybin_refine_factor, zbin_refine_factor),
                        max_cells_per_dim=max_cells_per_dim, copy_particles=copy_particles,
                        enable_min_sep_opt=enable_min_sep_opt, c_api_timer=c_api_timer, isa=isa,
                        fast_divide_and_NR_steps=fast_divide_and_NR_steps)
    w1 = None if weights1 is None else np.ascontiguousarray(weights1[0])
    w2 = None if weights2 is None else np.ascontiguousarray(weights2[0])
    r = _capi.call_DDsmu_mocks(_lib.load(), autocorr, cosmology, nthreads, mu_max, nmu_bins, binfile, RA1, DEC1, CZ1,
                               w1=w1, RA2=RA2, DEC2=DEC2, CZ2=CZ2, w2=w2, weight_type=weight_type, options=opt,
                               dtype=dtype)
    ns, nmu = r["npairs"].shape
    if r["npairs"].size == 0:  # empty particle set: an empty table, like the reference
        nmu = max(nmu, 1)
    res = np.zeros(ns * nmu, dtype=[("smin", np.float64), ("smax", np.float64), ("savg", np.float64),
                                    ("mumax", np.float64), ("npairs", np.uint64), ("weightavg", np.float64)])
    dmu = r["mu_max"] / nmu
    res["smin"] = np.repeat(r["rupp"][:-1], nmu)
    res["smax"] = np.repeat(r["rupp"][1:], nmu)
    res["mumax"] = np.tile((np.arange(nmu) + 1) * dmu, ns)
    res["savg"], res["npairs"], res["weightavg"] = r["ravg"].ravel(), r["npairs"].ravel(), r["weightavg"].ravel()
    return (res, r["api_time"]) if c_api_timer else res
