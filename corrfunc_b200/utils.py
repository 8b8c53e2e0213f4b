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
