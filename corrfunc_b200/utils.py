"""Host-side helpers shared by the Python wrappers (the parts of Corrfunc/utils.py the six wrappers
need: weights preparation, isa string parsing)."""
from __future__ import annotations

import numpy as np

_ISA = {"fallback": 0, "sse42": 6, "avx": 7, "avx2": 8, "avx512f": 9, "fastest": -1}


def translate_isa_string_to_enum(isa):
    """Accepted for API compatibility (Corrfunc/utils.py:464-518); the GPU path ignores it."""
    if isinstance(isa, (int, np.integer)):
        return int(isa)
    try:
        return _ISA[str(isa).lower()]
    except KeyError:
        raise ValueError("Do not know instruction type = %r. Valid values are %s" % (isa, sorted(_ISA)))


def process_weights(weights1, weights2, X1, X2, weight_type, autocorr):
    """Same contract as Corrfunc/utils.py:967-1026: scalars are broadcast, a missing set becomes
    ones for pair_product, arrays are returned with shape (n_weights, n_particles)."""
    if weight_type is None:
        return None, None

    def prep(w, x):
        if w is None:
            return None
        if np.isscalar(w) or (isinstance(w, np.ndarray) and w.ndim == 0):
            w = np.array(w, dtype=np.asarray(x).dtype)  # a scalar weight takes the particles' dtype (Corrfunc/utils.py:994-1003)
        w = np.atleast_1d(w)
        if w.shape[-1] == 1:
            w = np.tile(w, len(x))
        return np.atleast_2d(w)

    weights1 = prep(weights1, X1)
    if not autocorr:
        weights2 = prep(weights2, X2)
        if (weights1 is None) != (weights2 is None) and weight_type != "pair_product":
            raise ValueError("If using a weight_type other than 'pair_product', you must provide both weight arrays.")
        if weights1 is None and weights2 is not None:
            weights1 = np.ones((len(weights2), len(X1)), dtype=X1.dtype)
        if weights2 is None and weights1 is not None:
            weights2 = np.ones((len(weights1), len(X2)), dtype=X2.dtype)
    return weights1, weights2


def native_inputs(positions, weights1, weights2, X1, X2, weight_type, autocorr):
    """What every wrapper does with its arrays before the C call, in the reference's order (Corrfunc/theory/DD.py:222-247):
    weights are brought into shape first (scalars take the particles' dtype), every array is converted to the machine's byte
    order, and only then must they all share one dtype.  Returns (positions, weights1, weights2, dtype)."""
    weights1, weights2 = process_weights(weights1, weights2, X1, X2, weight_type, autocorr)
    positions = [convert_to_native_endian(a, warn=True) for a in positions]
    weights1 = convert_to_native_endian(weights1, warn=True)
    weights2 = convert_to_native_endian(weights2, warn=True)
    dtype = check_same_dtype(*positions, weights1, weights2)
    return positions, weights1, weights2, dtype


def check_same_dtype(*arrs):
    dt = None
    for a in arrs:
        if a is None:
            continue
        a = np.asarray(a)
        if a.dtype not in (np.float32, np.float64):
            raise TypeError("input arrays must be float32 or float64 (got %s)" % a.dtype)
        if dt is None:
            dt = a.dtype
        elif a.dtype != dt:
            raise TypeError("all input arrays must share one dtype (%s vs %s)" % (dt, a.dtype))
    return dt


# ------------------------------------------------------------------------------------------------
# Estimators: the step right after the pair counts.  Same names, argument order
# and error behaviour as Corrfunc/utils.py:27-165 (convert_3d_counts_to_cf) and :167-322
# (convert_rp_pi_counts_to_wp); pure numpy on the small per-bin arrays.
def _npairs_of(counts):
    """Accepts either a plain array of pair counts or the structured array a wrapper returned."""
    a = np.asarray(counts)
    if a.dtype.names and "npairs" in a.dtype.names:
        a = a["npairs"]
    return a.astype(np.float64)


def convert_3d_counts_to_cf(ND1, ND2, NR1, NR2, D1D2, D1R2, D2R1, R1R2, estimator="LS"):
    """Landy-Szalay correlation function from raw pair counts:
    ``(f1 f2 D1D2 - f1 D1R2 - f2 D2R1 + R1R2) / R1R2`` with ``f = NR / ND``; NaN where ``R1R2 == 0``."""
    if not ("LS" in estimator or "Landy" in estimator):
        raise ValueError("Only the Landy-Szalay estimator is supported. Pass estimator='LS'. "
                         "(Got estimator = {0})".format(estimator))
    dd, d1r2, d2r1, rr = (_npairs_of(a) for a in (D1D2, D1R2, D2R1, R1R2))
    if not (len(dd) == len(d1r2) == len(d2r1) == len(rr)):
        raise ValueError("Pair counts must have the same number of elements (same bins)")
    f1 = np.float64(NR1) / np.float64(ND1)
    f2 = np.float64(NR2) / np.float64(ND2)
    cf = np.full(len(dd), np.nan)
    ok = rr > 0
    cf[ok] = (f1 * f2 * dd[ok] - f1 * d1r2[ok] - f2 * d2r1[ok] + rr[ok]) / rr[ok]
    return cf


def convert_rp_pi_counts_to_wp(ND1, ND2, NR1, NR2, D1D2, D1R2, D2R1, R1R2, nrpbins, pimax, dpi=1.0, estimator="LS"):
    """Projected correlation function: ``wp(rp) = 2 dpi sum_pi xi(rp, pi)`` over the ``pimax / dpi`` line-of-sight
    bins of every rp bin; an rp bin with an empty RR pi-bin is NaN."""
    if dpi <= 0.0:
        raise ValueError("Binsize along the line of sight (dpi) = {0} must be positive".format(dpi))
    xi = convert_3d_counts_to_cf(ND1, ND2, NR1, NR2, D1D2, D1R2, D2R1, R1R2, estimator=estimator)
    npibins = len(xi) // nrpbins
    if npibins * nrpbins != len(xi):
        raise ValueError("Number of pi bins could not be calculated correctly: {0} bins are not a multiple of "
                         "the {1} rp bins".format(len(xi), nrpbins))
    if dpi * npibins != pimax:
        raise ValueError("Pimax = {0} should be equal to the product of npibins = {1} and dpi = {2}. "
                         "Check your binning scheme.".format(pimax, npibins, dpi))
    return 2.0 * dpi * xi.reshape(nrpbins, npibins).sum(axis=1)


# ------------------------------------------------------------------------------------------------
# The remaining public helpers of Corrfunc/utils.py, so that scripts importing them keep working.

def return_file_with_rbins(rbins):
    """(filename, delete_after_use) for a bin specification (Corrfunc/utils.py:324-381): an existing file name is
    returned as is; an array of edges is sorted and written as "low high" lines to a temporary file."""
    import os
    import tempfile

    if isinstance(rbins, str):
        if os.path.exists(rbins):
            return rbins, False
        raise IOError("Could not find file = `{0}` containing the bins".format(rbins))
    if len(rbins) >= 1:
        edges = sorted(rbins)
        with tempfile.NamedTemporaryFile(delete=False, mode="w") as f:
            for lo, hi in zip(edges[:-1], edges[1:]):
                f.write("{0} {1}\n".format(repr(float(lo)), repr(float(hi))))  # repr(float): parseable by sscanf("%lf")
            return f.name, True
    raise TypeError("Input `binfile` was not a valid array (>= 1 element).Num elements = {0}".format(len(rbins)))


def fix_cz(cz):
    """Redshifts passed where ``cz`` is expected (maximum below 10) are multiplied by the speed of light IN PLACE,
    as the reference does (Corrfunc/utils.py:384-418, SPEED_OF_LIGHT = 299800)."""
    try:
        input_dtype = cz.dtype
    except AttributeError:
        raise TypeError("Input cz array must be a numpy array")
    if cz.max() < 10.0:
        cz *= 299800.0
    return cz.astype(input_dtype)


def fix_ra_dec(ra, dec):
    """RA in [-180, 180] -> [0, 360] and DEC in [0, 180] -> [-90, 90], IN PLACE (Corrfunc/utils.py:421-461)."""
    try:
        input_dtype = ra.dtype
    except AttributeError:
        raise TypeError("Input RA array must be a numpy array")
    if ra is None or dec is None:
        raise ValueError("RA or DEC must be valid arrays")
    if ra.min() < 0.0:
        print("Warning: found negative RA values, wrapping into [0.0, 360.0]  range")
        ra += 180.0
    if dec.max() > 90.0:
        print("Warning: found DEC values more than 90.0; wrapping into [-90.0, 90.0] range")
        dec -= 90.0
    return ra.astype(input_dtype), dec.astype(input_dtype)


def compute_nbins(max_diff, binsize, refine_factor=1, max_nbins=None):
    """Number of cells of size >= ``binsize`` that span ``max_diff``, times ``refine_factor``, capped at
    ``max_nbins`` (Corrfunc/utils.py:521-596)."""
    if max_diff <= 0 or binsize <= 0:
        raise ValueError("Error: Invalid value for max_diff = {0} or binsize = {1}. Both must be positive"
                         .format(max_diff, binsize))
    if max_nbins is not None and max_nbins < 1:
        raise ValueError("Error: Invalid for the max. number of bins allowed = {0}.Max. nbins must be >= 1"
                         .format(max_nbins))
    if refine_factor < 1:
        raise ValueError("Error: Refine factor must be >=1. Found refine_factor = {0}".format(refine_factor))
    ngrid = max(1, int(max_diff / binsize)) * refine_factor
    if max_nbins:
        ngrid = min(int(max_nbins), ngrid)
    return ngrid


def gridlink_sphere(thetamax, ra_limits=None, dec_limits=None, link_in_ra=True, ra_refine_factor=1,
                    dec_refine_factor=1, max_ra_cells=100, max_dec_cells=200, return_num_ra_cells=False,
                    input_in_degrees=True):
    """The (DEC band, RA cell) lattice on the sphere for a maximum angular separation ``thetamax``
    (Corrfunc/utils.py:599-863): a structured array with fields ``dec_limit`` and ``ra_limit`` (two float64 each,
    radians), band by band; optionally also the number of RA cells per band."""
    from math import pi, radians

    if input_in_degrees:
        thetamax = radians(thetamax)
        ra_limits = [radians(x) for x in ra_limits] if ra_limits else ra_limits
        dec_limits = [radians(x) for x in dec_limits] if dec_limits else dec_limits
    if not ra_limits:
        ra_limits = [0.0, 2.0 * pi]
    if not dec_limits:
        dec_limits = [-0.5 * pi, 0.5 * pi]
    if dec_limits[0] >= dec_limits[1]:
        raise ValueError("Declination limits should be sorted in increasing order. However, dec_limits = [{0}, {1}] "
                         "is not".format(dec_limits[0], dec_limits[1]))
    if ra_limits[0] >= ra_limits[1]:
        raise ValueError("Declination limits should be sorted in increasing order. However, ra_limits = [{0}, {1}] "
                         "is not".format(ra_limits[0], ra_limits[1]))
    if dec_limits[0] < -0.5 * pi or dec_limits[1] > 0.5 * pi:
        raise ValueError("Valid range of values for declination are [-pi/2, +pi/2] deg. However, dec_limits = "
                         "[{0}, {1}] does not fall within that range".format(dec_limits[0], dec_limits[1]))
    if ra_limits[0] < 0.0 or ra_limits[1] > 2.0 * pi:
        raise ValueError("Valid range of values for declination are [0.0, 2*pi] deg. However, ra_limits = [{0}, {1}] "
                         "does not fall within that range".format(ra_limits[0], ra_limits[1]))
    dec_diff = abs(dec_limits[1] - dec_limits[0])
    ngrid_dec = compute_nbins(dec_diff, thetamax, refine_factor=dec_refine_factor, max_nbins=max_dec_cells)
    dec_binsize = dec_diff / ngrid_dec
    grid_dtype = np.dtype({"names": ["dec_limit", "ra_limit"], "formats": [(np.float64, (2,)), (np.float64, (2,))]})
    band = np.arange(ngrid_dec)
    if not link_in_ra:
        grid = np.zeros(ngrid_dec, dtype=grid_dtype)
        grid["dec_limit"][:, 0] = dec_limits[0] + band * dec_binsize
        grid["dec_limit"][:, 1] = dec_limits[0] + (band + 1) * dec_binsize
        grid["ra_limit"][:, 0] = ra_limits[0]
        grid["ra_limit"][:, 1] = ra_limits[1]
        return grid
    ra_diff = ra_limits[1] - ra_limits[0]
    sin_thetamax = np.sin(thetamax)  # the reference calls this sin_half_thetamax but takes the sine of thetamax
    num_ra_cells = np.full(ngrid_dec, ra_refine_factor, dtype=np.int64)
    for idec in range(ngrid_dec):
        dec_min = dec_limits[0] + idec * dec_binsize
        dec_max = dec_min + dec_binsize
        min_cos = min(np.cos(dec_min), np.cos(dec_max))
        if min_cos > 0:
            ratio = max(min(sin_thetamax / min_cos, 1.0), 0.0)
            ra_binsize = min(2.0 * np.arcsin(ratio), ra_diff)
            num_ra_cells[idec] = compute_nbins(ra_diff, ra_binsize, refine_factor=ra_refine_factor, max_nbins=max_ra_cells)
    grid = np.zeros(int(num_ra_cells.sum()), dtype=grid_dtype)
    ra_binsizes = ra_diff / num_ra_cells
    which = np.repeat(band, num_ra_cells)                                   # DEC band of every cell
    ira = np.arange(grid.size) - np.repeat(np.cumsum(num_ra_cells) - num_ra_cells, num_ra_cells)
    grid["dec_limit"][:, 0] = dec_limits[0] + dec_binsize * which
    grid["dec_limit"][:, 1] = dec_limits[0] + dec_binsize * (which + 1)
    grid["ra_limit"][:, 0] = ra_limits[0] + ra_binsizes[which] * ira
    grid["ra_limit"][:, 1] = ra_limits[0] + ra_binsizes[which] * (ira + 1)
    return (grid, num_ra_cells) if return_num_ra_cells else grid


def is_native_endian(array):
    """True when ``array`` is in the machine's byte order (None counts as native; Corrfunc/utils.py:924-964)."""
    if array is None:
        return True
    return np.asanyarray(array).dtype.isnative


def convert_to_native_endian(array, warn=False):
    """``array`` itself when it already has the machine's byte order, otherwise a byte-swapped native copy
    (Corrfunc/utils.py:866-921)."""
    if array is None:
        return array
    array = np.asanyarray(array)
    if array.dtype.isnative:
        return array
    if warn:
        import warnings

        warnings.warn("One or more input array has non-native endianness!  A copy will be made with the correct "
                      "endianness.")
    return array.astype(array.dtype.newbyteorder("="))


def sys_pipes():
    """The reference wraps its extension calls in ``wurlitzer.sys_pipes`` so that C-level output shows up in
    notebooks (Corrfunc/utils.py:1030-1068); this library writes its messages to the process's stderr, so the
    context manager has nothing to redirect."""
    import contextlib

    return contextlib.nullcontext()
