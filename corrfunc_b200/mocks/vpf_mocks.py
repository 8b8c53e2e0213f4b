"""``Corrfunc.mocks.vpf_mocks`` look-alike (reference: Corrfunc/mocks/vpf_mocks.py:17-330)."""
from __future__ import annotations

import numpy as np

from .. import _capi, _lib
from ..utils import check_same_dtype, convert_to_native_endian, translate_isa_string_to_enum


def vpf_mocks(rmax, nbins, nspheres, numpN, threshold_ngb, centers_file, cosmology, RA, DEC, CZ, RAND_RA, RAND_DEC,
              RAND_CZ, verbose=False, is_comoving_dist=False, xbin_refine_factor=1, ybin_refine_factor=1,
              zbin_refine_factor=1, max_cells_per_dim=100, copy_particles=True, c_api_timer=False, isa="fastest"):
    """Counts-in-spheres on a survey catalogue: the probability pN that a sphere of radius r holds exactly N galaxies,
    for ``nbins`` radii up to ``rmax`` and N < ``numpN``.  Sphere centres are read from ``centers_file`` when it holds
    enough of them, otherwise placed on the randoms (and the file is rewritten), as the reference does.  Returns a
    structured array (rmax, pN[numpN]) [and the C call's wall time when ``c_api_timer``]."""
    translate_isa_string_to_enum(isa)
    RA, DEC, CZ, RAND_RA, RAND_DEC, RAND_CZ = [convert_to_native_endian(a, warn=True)
                                               for a in (RA, DEC, CZ, RAND_RA, RAND_DEC, RAND_CZ)]
    dtype = check_same_dtype(RA, DEC, CZ, RAND_RA, RAND_DEC, RAND_CZ)
    refine = (xbin_refine_factor, ybin_refine_factor, zbin_refine_factor)
    opt = _capi.default_options(dtype, verbose=verbose, periodic=False, bin_refine_factors=refine,
                                max_cells_per_dim=max_cells_per_dim, copy_particles=copy_particles,
                                c_api_timer=c_api_timer, isa=-1, custom_refine=tuple(int(r) for r in refine) != (1, 1, 1),
                                is_comoving_dist=is_comoving_dist)
    r = _capi.call_vpf_mocks(_lib.load(), rmax, nbins, nspheres, numpN, threshold_ngb, centers_file, cosmology, RA, DEC,
                             CZ, RAND_RA=RAND_RA, RAND_DEC=RAND_DEC, RAND_CZ=RAND_CZ, options=opt, dtype=dtype)
    res = np.zeros(r["nbin"], dtype=[("rmax", np.float64), ("pN", (np.float64, numpN))])
    rstep = rmax / float(nbins)  # _countpairs_mocks.c:2230-2232
    res["rmax"] = (np.arange(r["nbin"]) + 1) * rstep
    res["pN"] = r["pN"] if numpN > 1 else r["pN"][:, 0]
    return (res, r["api_time"]) if c_api_timer else res
