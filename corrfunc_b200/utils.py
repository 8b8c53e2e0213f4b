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
        if isinstance(w, float):
            w = np.array(w, dtype=x.dtype)
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
