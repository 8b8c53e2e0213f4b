"""DD / DR / RR in one GPU context -- the step right after the pair counts in the reference's workflow
(``Corrfunc/utils.py:27-165`` ``convert_3d_counts_to_cf``, ``:167-322`` ``convert_rp_pi_counts_to_wp``; their docstring
examples call ``DD`` three times over a data and a random catalogue).

Three separate calls upload and sort each catalogue twice.  Inside ``shared_catalogs()`` the library keeps what it was
given under the same array pointers (``corrfunc_b200_catalog_cache``): D and R cross PCIe once, stay resident, and are
sorted again only when the lattice of the next count differs.  The counts are exactly those of three separate calls."""
from __future__ import annotations

from contextlib import contextmanager

import numpy as np

from . import _lib
from .theory import DD, DDrppi, DDsmu
from .utils import convert_3d_counts_to_cf, convert_rp_pi_counts_to_wp


@contextmanager
def shared_catalogs():
    """While active, arrays passed again under the same pointers are taken to be unchanged (do not modify them)."""
    lib = _lib.load()
    lib.corrfunc_b200_catalog_cache(1)
    try:
        yield
    finally:
        lib.corrfunc_b200_catalog_cache(0)


def _native(a, dtype):
    return None if a is None else np.ascontiguousarray(a, dtype=dtype)


def DD_DR_RR(nthreads, binfile, X, Y, Z, RX, RY, RZ, weights=None, rweights=None, weight_type=None, periodic=True,
             boxsize=None, stat="DD", pimax=None, mu_max=None, nmu_bins=None, **kwargs):
    """Auto-correlation of the data, data x randoms and auto-correlation of the randoms, with both catalogues uploaded
    once.  ``stat``: "DD" (3-D r bins), "DDrppi" (needs pimax) or "DDsmu" (needs mu_max, nmu_bins); remaining keyword
    arguments go to the counting function.  Returns the three structured arrays (DD, DR, RR)."""
    dtype = np.asarray(X).dtype
    # one native, contiguous copy of every array for all three calls: the cache keys on the pointers
    D = [_native(a, dtype) for a in (X, Y, Z)]
    R = [_native(a, dtype) for a in (RX, RY, RZ)]
    w, rw = _native(weights, dtype), _native(rweights, dtype)
    common = dict(periodic=periodic, boxsize=boxsize, weight_type=weight_type, **kwargs)

    def count(autocorr, A, wa, B=None, wb=None):
        kw = dict(common, weights1=wa)
        if B is not None:
            kw.update(X2=B[0], Y2=B[1], Z2=B[2], weights2=wb)
        if stat == "DD":
            return DD(autocorr, nthreads, binfile, A[0], A[1], A[2], **kw)
        if stat == "DDrppi":
            return DDrppi(autocorr, nthreads, pimax, binfile, A[0], A[1], A[2], **kw)
        if stat == "DDsmu":
            return DDsmu(autocorr, nthreads, binfile, mu_max, nmu_bins, A[0], A[1], A[2], **kw)
        raise ValueError("stat must be DD, DDrppi or DDsmu")

    with shared_catalogs():
        dd = count(1, D, w)
        dr = count(0, D, w, R, rw)
        rr = count(1, R, rw)
    return dd, dr, rr


def xi_from_catalogs(nthreads, binfile, X, Y, Z, RX, RY, RZ, estimator="LS", **kwargs):
    """Landy-Szalay xi(r) of a data catalogue against randoms: DD, DR, RR in one context, then
    ``convert_3d_counts_to_cf`` (Corrfunc/utils.py:27-165)."""
    dd, dr, rr = DD_DR_RR(nthreads, binfile, X, Y, Z, RX, RY, RZ, stat="DD", **kwargs)
    nd, nr = len(X), len(RX)
    return convert_3d_counts_to_cf(nd, nd, nr, nr, dd, dr, dr, rr, estimator=estimator)


def wp_from_catalogs(nthreads, pimax, binfile, X, Y, Z, RX, RY, RZ, estimator="LS", **kwargs):
    """wp(rp) from DD(rp, pi), DR, RR in one context and ``convert_rp_pi_counts_to_wp`` (Corrfunc/utils.py:167-322)."""
    dd, dr, rr = DD_DR_RR(nthreads, binfile, X, Y, Z, RX, RY, RZ, stat="DDrppi", pimax=pimax, **kwargs)
    nd, nr = len(X), len(RX)
    nrpbins = len(np.unique(dd["rmin"]))
    return convert_rp_pi_counts_to_wp(nd, nd, nr, nr, dd, dr, dr, rr, nrpbins, pimax, estimator=estimator)
