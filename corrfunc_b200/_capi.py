"""ctypes view of the Corrfunc C ABI (struct layouts + the six entry points).

The layouts mirror ``include/corrfunc_b200_defs.h`` / ``include/countpairs*.h`` (which in turn keep
the binary layout of the reference's ``utils/defs.h:53-156,353-402`` and result structs).  The
callers in this module work on *any* library exporting that ABI: the product library
``libcorrfunc_b200.so`` (see :mod:`corrfunc_b200._lib`) and -- in the tests only -- the unmodified
reference built by the test harness.
"""
from __future__ import annotations

import ctypes as C
import os
import tempfile
from contextlib import contextmanager

import numpy as np

OPTIONS_HEADER_SIZE = 1024
BOXSIZE_NOTGIVEN = -2.0
MAX_NUM_WEIGHTS = 10
WEIGHT_NONE = -42
WEIGHT_PAIR_PRODUCT = 0


class _Boxsize(C.Union):
    _fields_ = [("boxsize", C.c_double), ("boxsize_x", C.c_double)]


class _BinFlags(C.Union):
    _fields_ = [("binning_flags", C.c_uint32), ("bin_masks", C.c_uint8 * 4)]


class ConfigOptions(C.Structure):
    _anonymous_ = ("_b", "_f")
    _fields_ = [
        ("_b", _Boxsize),
        ("boxsize_y", C.c_double),
        ("boxsize_z", C.c_double),
        ("OMEGA_M", C.c_double),
        ("OMEGA_B", C.c_double),
        ("OMEGA_L", C.c_double),
        ("HUBBLE", C.c_double),
        ("LITTLE_H", C.c_double),
        ("SIGMA_8", C.c_double),
        ("NS", C.c_double),
        ("c_api_time", C.c_double),
        ("cell_timings", C.c_void_p),
        ("totncells_timings", C.c_int64),
        ("float_type", C.c_size_t),
        ("instruction_set", C.c_int32),
        ("version", C.c_char * 32),
        ("verbose", C.c_uint8),
        ("c_api_timer", C.c_uint8),
        ("c_cell_timer", C.c_uint8),
        ("need_avg_sep", C.c_uint8),
        ("autocorr", C.c_uint8),
        ("periodic", C.c_uint8),
        ("sort_on_z", C.c_uint8),
        ("is_comoving_dist", C.c_uint8),
        ("link_in_dec", C.c_uint8),
        ("link_in_ra", C.c_uint8),
        ("fast_divide_and_NR_steps", C.c_uint8),
        ("fast_acos", C.c_uint8),
        ("enable_min_sep_opt", C.c_uint8),
        ("bin_refine_factors", C.c_int8 * 3),
        ("max_cells_per_dim", C.c_uint16),
        ("copy_particles", C.c_uint8),
        ("use_heap_sort", C.c_uint8),
        ("_f", _BinFlags),
        ("reserved", C.c_uint8 * 1),  # resized below
    ]


def _pad_struct(cls, size):
    # ctypes has no flexible "fill to N bytes"; compute the tail once and rebuild the field list
    base = [f for f in cls._fields_ if f[0] != "reserved"]

    class _Probe(C.Structure):
        _anonymous_ = getattr(cls, "_anonymous_", ())
        _fields_ = base

    tail = size - C.sizeof(_Probe)
    assert tail > 0

    class _Final(C.Structure):
        _anonymous_ = getattr(cls, "_anonymous_", ())
        _fields_ = base + [("reserved", C.c_uint8 * tail)]

    _Final.__name__ = cls.__name__
    assert C.sizeof(_Final) == size, (C.sizeof(_Final), size)
    return _Final


ConfigOptions = _pad_struct(ConfigOptions, OPTIONS_HEADER_SIZE)


class WeightStruct(C.Structure):
    _fields_ = [("weights", C.c_void_p * MAX_NUM_WEIGHTS), ("num_weights", C.c_int64)]


class ExtraOptions(C.Structure):
    _fields_ = [
        ("weights0", WeightStruct),
        ("weights1", WeightStruct),
        ("weight_method", C.c_int),
        ("reserved", C.c_uint8 * 1),
    ]


ExtraOptions = _pad_struct(ExtraOptions, OPTIONS_HEADER_SIZE)

_u64p = C.POINTER(C.c_uint64)
_f64p = C.POINTER(C.c_double)


class ResultsDD(C.Structure):
    _fields_ = [("npairs", _u64p), ("rupp", _f64p), ("rpavg", _f64p), ("weightavg", _f64p), ("nbin", C.c_int)]


class ResultsRpPi(C.Structure):
    _fields_ = [("npairs", _u64p), ("rupp", _f64p), ("rpavg", _f64p), ("weightavg", _f64p),
                ("pimax", C.c_double), ("nbin", C.c_int), ("npibin", C.c_int)]


class ResultsSMu(C.Structure):
    _fields_ = [("npairs", _u64p), ("supp", _f64p), ("savg", _f64p), ("mu_max", C.c_double),
                ("mu_min", C.c_double), ("weightavg", _f64p), ("nsbin", C.c_int), ("nmu_bins", C.c_int)]


class ResultsMocksRpPi(ResultsRpPi):  # results_countpairs_mocks (countpairs_rp_pi_mocks.h:18-26): same layout
    pass


class ResultsMocksSMu(ResultsSMu):  # results_countpairs_mocks_s_mu (countpairs_s_mu_mocks.h:19-28): same layout
    pass


class ResultsVpfMocks(C.Structure):  # results_countspheres_mocks (mocks/vpf_mocks/countspheres_mocks.h:20-26)
    _fields_ = [("pN", C.POINTER(C.POINTER(C.c_double))), ("rmax", C.c_double), ("nbin", C.c_int), ("nc", C.c_int),
                ("num_pN", C.c_int)]


class ResultsWp(C.Structure):
    _fields_ = [("npairs", _u64p), ("wp", _f64p), ("rupp", _f64p), ("rpavg", _f64p), ("weightavg", _f64p),
                ("pimax", C.c_double), ("nbin", C.c_int)]


class ResultsXi(C.Structure):
    _fields_ = [("npairs", _u64p), ("xi", _f64p), ("rupp", _f64p), ("ravg", _f64p), ("weightavg", _f64p),
                ("nbin", C.c_int)]


class ResultsTheta(C.Structure):
    _fields_ = [("npairs", _u64p), ("theta_upp", _f64p), ("theta_avg", _f64p), ("weightavg", _f64p),
                ("nbin", C.c_int)]


def default_options(dtype, *, verbose=False, periodic=True, need_avg_sep=False, boxsize=None,
                    bin_refine_factors=(2, 2, 1), max_cells_per_dim=100, copy_particles=True,
                    enable_min_sep_opt=True, c_api_timer=False, isa=-1, link_in_dec=True, link_in_ra=True,
                    fast_acos=False, custom_refine=False, is_comoving_dist=False) -> ConfigOptions:
    """Python twin of ``get_config_options()`` plus the kwargs the reference's extension sets
    (theory/python_bindings/_countpairs.c:1153-1260)."""
    o = ConfigOptions()
    # the reference stringifies a quoted macro, so its version string carries the quote characters
    o.version = b'"2.5.3"'
    o.boxsize_x = o.boxsize_y = o.boxsize_z = BOXSIZE_NOTGIVEN
    if boxsize is not None:
        bx = np.atleast_1d(np.asarray(boxsize, dtype=np.float64))
        if bx.size == 1:
            o.boxsize_x = float(bx[0]); o.boxsize_y = float(bx[0]); o.boxsize_z = float(bx[0])
        else:
            o.boxsize_x, o.boxsize_y, o.boxsize_z = (float(b) for b in bx[:3])
    o.float_type = np.dtype(dtype).itemsize
    o.instruction_set = int(isa)
    o.verbose = int(bool(verbose))
    o.c_api_timer = int(bool(c_api_timer))
    o.need_avg_sep = int(bool(need_avg_sep))
    o.periodic = int(bool(periodic))
    o.link_in_dec = int(bool(link_in_dec))
    o.link_in_ra = int(bool(link_in_ra))
    o.fast_acos = int(bool(fast_acos))
    o.is_comoving_dist = int(bool(is_comoving_dist))
    o.enable_min_sep_opt = int(bool(enable_min_sep_opt))
    o.copy_particles = int(bool(copy_particles))
    for i in range(3):
        o.bin_refine_factors[i] = int(bin_refine_factors[i])
    o.max_cells_per_dim = int(max_cells_per_dim)
    o.binning_flags = 1 if custom_refine else 0
    return o


def make_extra(weights1, weights2, weight_type, dtype):
    """Fill ``struct extra_options``; returns (extra, keepalive list)."""
    e = ExtraOptions()
    keep = []
    if weight_type is None or weight_type == "":
        e.weight_method = WEIGHT_NONE
        return e, keep
    if weight_type not in ("pair_product", "p"):
        raise ValueError("unknown weight_type %r" % (weight_type,))
    e.weight_method = WEIGHT_PAIR_PRODUCT
    for ws, arr in ((e.weights0, weights1), (e.weights1, weights2)):
        if arr is None:
            continue
        a = np.ascontiguousarray(np.atleast_2d(arr), dtype=dtype)
        keep.append(a)
        ws.num_weights = 1
        ws.weights[0] = a[0].ctypes.data
    return e, keep


@contextmanager
def binfile_for(bins):
    """The reference's Python layer writes array bins to a temp 'lo hi' file
    (Corrfunc/utils.py:324-381); a string is taken as a path."""
    if isinstance(bins, (str, bytes, os.PathLike)):
        yield os.fsencode(bins)
        return
    edges = np.sort(np.asarray(bins, dtype=np.float64))
    if edges.size < 2:
        raise ValueError("need at least two bin edges")
    with tempfile.NamedTemporaryFile("w", suffix=".bins", delete=False) as f:
        for lo, hi in zip(edges[:-1], edges[1:]):
            f.write("%s %s\n" % (repr(float(lo)), repr(float(hi))))
        name = f.name
    try:
        yield os.fsencode(name)
    finally:
        os.unlink(name)


def _ptr(a):
    return C.c_void_p(a.ctypes.data) if a is not None else C.c_void_p(None)


def _arr(p, n, dtype):
    if not p or n <= 0:
        return np.zeros(0, dtype=dtype)
    return np.ctypeslib.as_array(p, shape=(n,)).astype(dtype, copy=True)


def _prep(arrs, dtype):
    out = []
    for a in arrs:
        out.append(None if a is None else np.ascontiguousarray(a, dtype=dtype))
    return out


def _declare(lib):
    if getattr(lib, "_cf_declared", False):
        return
    vp, i64, ci, cd, cs = C.c_void_p, C.c_int64, C.c_int, C.c_double, C.c_char_p
    lib.countpairs.argtypes = [i64, vp, vp, vp, i64, vp, vp, vp, ci, ci, cs, C.POINTER(ResultsDD),
                               C.POINTER(ConfigOptions), C.POINTER(ExtraOptions)]
    lib.countpairs_rp_pi.argtypes = [i64, vp, vp, vp, i64, vp, vp, vp, ci, ci, cs, cd, C.POINTER(ResultsRpPi),
                                     C.POINTER(ConfigOptions), C.POINTER(ExtraOptions)]
    lib.countpairs_s_mu.argtypes = [i64, vp, vp, vp, i64, vp, vp, vp, ci, ci, cs, cd, ci, C.POINTER(ResultsSMu),
                                    C.POINTER(ConfigOptions), C.POINTER(ExtraOptions)]
    lib.countpairs_wp.argtypes = [i64, vp, vp, vp, cd, ci, cs, cd, C.POINTER(ResultsWp),
                                  C.POINTER(ConfigOptions), C.POINTER(ExtraOptions)]
    lib.countpairs_xi.argtypes = [i64, vp, vp, vp, cd, ci, cs, C.POINTER(ResultsXi),
                                  C.POINTER(ConfigOptions), C.POINTER(ExtraOptions)]
    lib.countpairs_theta_mocks.argtypes = [i64, vp, vp, i64, vp, vp, ci, ci, cs, C.POINTER(ResultsTheta),
                                           C.POINTER(ConfigOptions), C.POINTER(ExtraOptions)]
    for f in ("countpairs", "countpairs_rp_pi", "countpairs_s_mu", "countpairs_wp", "countpairs_xi",
              "countpairs_theta_mocks"):
        getattr(lib, f).restype = ci
    for f, t in (("free_results", ResultsDD), ("free_results_rp_pi", ResultsRpPi),
                 ("free_results_s_mu", ResultsSMu), ("free_results_wp", ResultsWp),
                 ("free_results_xi", ResultsXi), ("free_results_countpairs_theta", ResultsTheta)):
        getattr(lib, f).argtypes = [C.POINTER(t)]
        getattr(lib, f).restype = None
    # mocks/DDrppi_mocks, mocks/DDsmu_mocks (absent from reference builds without them)
    if hasattr(lib, "countpairs_mocks"):
        lib.countpairs_mocks.argtypes = [i64, vp, vp, vp, i64, vp, vp, vp, ci, ci, cs, cd, ci,
                                         C.POINTER(ResultsMocksRpPi), C.POINTER(ConfigOptions), C.POINTER(ExtraOptions)]
        lib.countpairs_mocks_s_mu.argtypes = [i64, vp, vp, vp, i64, vp, vp, vp, ci, ci, cs, cd, ci, ci,
                                              C.POINTER(ResultsMocksSMu), C.POINTER(ConfigOptions), C.POINTER(ExtraOptions)]
        lib.countpairs_mocks.restype = lib.countpairs_mocks_s_mu.restype = ci
        lib.free_results_mocks.argtypes = [C.POINTER(ResultsMocksRpPi)]
        lib.free_results_mocks_s_mu.argtypes = [C.POINTER(ResultsMocksSMu)]
        lib.free_results_mocks.restype = lib.free_results_mocks_s_mu.restype = None
    if hasattr(lib, "countspheres"):  # theory/vpf
        lib.countspheres.argtypes = [i64, vp, vp, vp, cd, ci, ci, ci, C.c_ulong, C.POINTER(ResultsVpfMocks),
                                     C.POINTER(ConfigOptions), C.POINTER(ExtraOptions)]
        lib.countspheres.restype = ci
        lib.free_results_countspheres.argtypes = [C.POINTER(ResultsVpfMocks)]
        lib.free_results_countspheres.restype = None
    if hasattr(lib, "countspheres_mocks"):  # mocks/vpf_mocks
        lib.countspheres_mocks.argtypes = [i64, vp, vp, vp, i64, vp, vp, vp, ci, cd, ci, ci, ci, cs, ci,
                                           C.POINTER(ResultsVpfMocks), C.POINTER(ConfigOptions), C.POINTER(ExtraOptions)]
        lib.countspheres_mocks.restype = ci
        lib.free_results_countspheres_mocks.argtypes = [C.POINTER(ResultsVpfMocks)]
        lib.free_results_countspheres_mocks.restype = None
    lib._cf_declared = True


EXPORTED_SYMBOLS = ("countpairs", "free_results", "countpairs_rp_pi", "free_results_rp_pi", "countpairs_s_mu",
                    "free_results_s_mu", "countpairs_wp", "free_results_wp", "countpairs_xi", "free_results_xi",
                    "countpairs_theta_mocks", "free_results_countpairs_theta", "countpairs_mocks", "free_results_mocks",
                    "countpairs_mocks_s_mu", "free_results_mocks_s_mu", "countspheres_mocks",
                    "free_results_countspheres_mocks", "countspheres", "free_results_countspheres", "countpairs_mocks_float",
                    "countpairs_mocks_double", "countpairs_mocks_s_mu_float", "countpairs_mocks_s_mu_double") + tuple(
    "%s_%s" % (f, t) for f in ("countpairs", "countpairs_rp_pi", "countpairs_s_mu", "countpairs_wp", "countpairs_xi",
                               "countpairs_theta_mocks") for t in ("float", "double"))


def call_DD(lib, autocorr, nthreads, bins, X1, Y1, Z1, w1=None, X2=None, Y2=None, Z2=None, w2=None,
            weight_type=None, options=None, dtype=None):
    _declare(lib)
    dtype = np.dtype(dtype or np.asarray(X1).dtype)
    X1, Y1, Z1, X2, Y2, Z2 = _prep((X1, Y1, Z1, X2, Y2, Z2), dtype)
    if autocorr:
        X2, Y2, Z2 = X1, Y1, Z1
    extra, keep = make_extra(w1, w2 if not autocorr else w1, weight_type, dtype)
    res = ResultsDD()
    n2 = 0 if X2 is None else X2.size
    with binfile_for(bins) as bf:
        st = lib.countpairs(X1.size, _ptr(X1), _ptr(Y1), _ptr(Z1), n2, _ptr(X2), _ptr(Y2), _ptr(Z2),
                            int(nthreads), int(autocorr), bf, C.byref(res), C.byref(options), C.byref(extra))
    if st != 0:
        raise RuntimeError("countpairs returned %d" % st)
    n = res.nbin
    out = dict(npairs=_arr(res.npairs, n, np.uint64)[1:], rupp=_arr(res.rupp, n, np.float64),
               ravg=_arr(res.rpavg, n, np.float64)[1:], weightavg=_arr(res.weightavg, n, np.float64)[1:],
               api_time=options.c_api_time)
    lib.free_results(C.byref(res))
    return out


def call_DDrppi(lib, autocorr, nthreads, pimax, bins, X1, Y1, Z1, w1=None, X2=None, Y2=None, Z2=None, w2=None,
                weight_type=None, options=None, dtype=None):
    _declare(lib)
    dtype = np.dtype(dtype or np.asarray(X1).dtype)
    X1, Y1, Z1, X2, Y2, Z2 = _prep((X1, Y1, Z1, X2, Y2, Z2), dtype)
    if autocorr:
        X2, Y2, Z2 = X1, Y1, Z1
    extra, keep = make_extra(w1, w2 if not autocorr else w1, weight_type, dtype)
    res = ResultsRpPi()
    n2 = 0 if X2 is None else X2.size
    with binfile_for(bins) as bf:
        st = lib.countpairs_rp_pi(X1.size, _ptr(X1), _ptr(Y1), _ptr(Z1), n2, _ptr(X2), _ptr(Y2), _ptr(Z2),
                                  int(nthreads), int(autocorr), bf, float(pimax), C.byref(res),
                                  C.byref(options), C.byref(extra))
    if st != 0:
        raise RuntimeError("countpairs_rp_pi returned %d" % st)
    nb, npi = res.nbin, res.npibin
    tot = (nb + 1) * (npi + 1)

    def grid(p, dt):
        a = _arr(p, tot, dt).reshape(nb + 1, npi + 1)
        return a[1:nb, :npi].copy()

    out = dict(npairs=grid(res.npairs, np.uint64), rupp=_arr(res.rupp, nb, np.float64),
               ravg=grid(res.rpavg, np.float64), weightavg=grid(res.weightavg, np.float64), npibin=npi,
               pimax=res.pimax, api_time=options.c_api_time)
    lib.free_results_rp_pi(C.byref(res))
    return out


def call_DDsmu(lib, autocorr, nthreads, bins, mu_max, nmu_bins, X1, Y1, Z1, w1=None, X2=None, Y2=None, Z2=None,
               w2=None, weight_type=None, options=None, dtype=None):
    _declare(lib)
    dtype = np.dtype(dtype or np.asarray(X1).dtype)
    X1, Y1, Z1, X2, Y2, Z2 = _prep((X1, Y1, Z1, X2, Y2, Z2), dtype)
    if autocorr:
        X2, Y2, Z2 = X1, Y1, Z1
    extra, keep = make_extra(w1, w2 if not autocorr else w1, weight_type, dtype)
    res = ResultsSMu()
    n2 = 0 if X2 is None else X2.size
    with binfile_for(bins) as bf:
        st = lib.countpairs_s_mu(X1.size, _ptr(X1), _ptr(Y1), _ptr(Z1), n2, _ptr(X2), _ptr(Y2), _ptr(Z2),
                                 int(nthreads), int(autocorr), bf, float(mu_max), int(nmu_bins), C.byref(res),
                                 C.byref(options), C.byref(extra))
    if st != 0:
        raise RuntimeError("countpairs_s_mu returned %d" % st)
    nb, nmu = res.nsbin, res.nmu_bins
    tot = (nb + 1) * (nmu + 1)

    def grid(p, dt):
        a = _arr(p, tot, dt).reshape(nb + 1, nmu + 1)
        return a[1:nb, :nmu].copy()

    out = dict(npairs=grid(res.npairs, np.uint64), rupp=_arr(res.supp, nb, np.float64),
               ravg=grid(res.savg, np.float64), weightavg=grid(res.weightavg, np.float64), nmu_bins=nmu,
               mu_max=res.mu_max, api_time=options.c_api_time)
    lib.free_results_s_mu(C.byref(res))
    return out


def call_DDrppi_mocks(lib, autocorr, cosmology, nthreads, pimax, bins, RA1, DEC1, CZ1, w1=None, RA2=None, DEC2=None,
                      CZ2=None, w2=None, weight_type=None, options=None, dtype=None):
    """countpairs_mocks (mocks/DDrppi_mocks/countpairs_rp_pi_mocks.h:28-37).  The C routine may shift RA/DEC in
    place (check_ra_dec_cz) -> private copies."""
    _declare(lib)
    dtype = np.dtype(dtype or np.asarray(RA1).dtype)
    RA1, DEC1, CZ1, RA2, DEC2, CZ2 = (None if a is None else np.array(a, dtype=dtype, order="C", copy=True)
                                      for a in (RA1, DEC1, CZ1, RA2, DEC2, CZ2))
    if autocorr:
        RA2, DEC2, CZ2 = RA1, DEC1, CZ1
    extra, keep = make_extra(w1, w2 if not autocorr else w1, weight_type, dtype)
    res = ResultsMocksRpPi()
    n2 = 0 if RA2 is None else RA2.size
    with binfile_for(bins) as bf:
        st = lib.countpairs_mocks(RA1.size, _ptr(RA1), _ptr(DEC1), _ptr(CZ1), n2, _ptr(RA2), _ptr(DEC2), _ptr(CZ2),
                                  int(nthreads), int(autocorr), bf, float(pimax), int(cosmology), C.byref(res),
                                  C.byref(options), C.byref(extra))
    if st != 0:
        raise RuntimeError("countpairs_mocks returned %d" % st)
    if not res.npairs:  # an empty particle set: EXIT_SUCCESS with the results untouched (rp_pi_mocks_impl:243-245)
        e = np.zeros((0, 0))
        return dict(npairs=e.astype(np.uint64), rupp=np.zeros(0), ravg=e, weightavg=e, npibin=0, pimax=float(pimax),
                    api_time=options.c_api_time)
    nb, npi = res.nbin, res.npibin
    tot = (nb + 1) * (npi + 1)

    def grid(p, dt):
        a = _arr(p, tot, dt).reshape(nb + 1, npi + 1)
        return a[1:nb, :npi].copy()

    out = dict(npairs=grid(res.npairs, np.uint64), rupp=_arr(res.rupp, nb, np.float64),
               ravg=grid(res.rpavg, np.float64), weightavg=grid(res.weightavg, np.float64), npibin=npi,
               pimax=res.pimax, api_time=options.c_api_time)
    lib.free_results_mocks(C.byref(res))
    return out


def call_DDsmu_mocks(lib, autocorr, cosmology, nthreads, mu_max, nmu_bins, bins, RA1, DEC1, CZ1, w1=None, RA2=None,
                     DEC2=None, CZ2=None, w2=None, weight_type=None, options=None, dtype=None):
    """countpairs_mocks_s_mu (mocks/DDsmu_mocks/countpairs_s_mu_mocks.h:30-41)."""
    _declare(lib)
    dtype = np.dtype(dtype or np.asarray(RA1).dtype)
    RA1, DEC1, CZ1, RA2, DEC2, CZ2 = (None if a is None else np.array(a, dtype=dtype, order="C", copy=True)
                                      for a in (RA1, DEC1, CZ1, RA2, DEC2, CZ2))
    if autocorr:
        RA2, DEC2, CZ2 = RA1, DEC1, CZ1
    extra, keep = make_extra(w1, w2 if not autocorr else w1, weight_type, dtype)
    res = ResultsMocksSMu()
    n2 = 0 if RA2 is None else RA2.size
    with binfile_for(bins) as bf:
        st = lib.countpairs_mocks_s_mu(RA1.size, _ptr(RA1), _ptr(DEC1), _ptr(CZ1), n2, _ptr(RA2), _ptr(DEC2),
                                       _ptr(CZ2), int(nthreads), int(autocorr), bf, float(mu_max), int(nmu_bins),
                                       int(cosmology), C.byref(res), C.byref(options), C.byref(extra))
    if st != 0:
        raise RuntimeError("countpairs_mocks_s_mu returned %d" % st)
    if not res.npairs:  # an empty particle set: EXIT_SUCCESS with the results untouched
        e = np.zeros((0, 0))
        return dict(npairs=e.astype(np.uint64), rupp=np.zeros(0), ravg=e, weightavg=e, nmu_bins=0, mu_max=float(mu_max),
                    api_time=options.c_api_time)
    nb, nmu = res.nsbin, res.nmu_bins
    tot = (nb + 1) * (nmu + 1)

    def grid(p, dt):
        a = _arr(p, tot, dt).reshape(nb + 1, nmu + 1)
        return a[1:nb, :nmu].copy()

    out = dict(npairs=grid(res.npairs, np.uint64), rupp=_arr(res.supp, nb, np.float64),
               ravg=grid(res.savg, np.float64), weightavg=grid(res.weightavg, np.float64), nmu_bins=nmu,
               mu_max=res.mu_max, api_time=options.c_api_time)
    lib.free_results_mocks_s_mu(C.byref(res))
    return out


def call_vpf(lib, rmax, nbin, nc, num_pN, seed, X, Y, Z, options=None, dtype=None):
    """countspheres (theory/vpf/countspheres.h:28-35; results_countspheres has the layout of ResultsVpfMocks)."""
    _declare(lib)
    dtype = np.dtype(dtype or np.asarray(X).dtype)
    X, Y, Z = _prep((X, Y, Z), dtype)
    extra, keep = make_extra(None, None, None, dtype)
    res = ResultsVpfMocks()
    st = lib.countspheres(X.size, _ptr(X), _ptr(Y), _ptr(Z), float(rmax), int(nbin), int(nc), int(num_pN), int(seed),
                          C.byref(res), C.byref(options), C.byref(extra))
    if st != 0:
        raise RuntimeError("countspheres returned %d" % st)
    pN = np.array([[res.pN[i][j] for j in range(res.num_pN)] for i in range(res.nbin)], dtype=np.float64)
    out = dict(pN=pN, rmax=res.rmax, nbin=res.nbin, nc=res.nc, api_time=options.c_api_time)
    lib.free_results_countspheres(C.byref(res))
    return out


def call_vpf_mocks(lib, rmax, nbin, nc, num_pN, threshold_neighbors, centers_file, cosmology, RA, DEC, CZ,
                   RAND_RA=None, RAND_DEC=None, RAND_CZ=None, options=None, dtype=None):
    """countspheres_mocks (mocks/vpf_mocks/countspheres_mocks.h:28-37): counts-in-spheres -> pN[nbin][num_pN]."""
    _declare(lib)
    dtype = np.dtype(dtype or np.asarray(RA).dtype)
    arrs = [None if a is None else np.array(a, dtype=dtype, order="C", copy=True)
            for a in (RA, DEC, CZ, RAND_RA, RAND_DEC, RAND_CZ)]
    RA, DEC, CZ, RR, RD, RC = arrs
    if RR is None:
        RR, RD, RC = RA, DEC, CZ
    extra, keep = make_extra(None, None, None, dtype)
    res = ResultsVpfMocks()
    st = lib.countspheres_mocks(RA.size, _ptr(RA), _ptr(DEC), _ptr(CZ), RR.size, _ptr(RR), _ptr(RD), _ptr(RC),
                                int(threshold_neighbors), float(rmax), int(nbin), int(nc), int(num_pN),
                                os.fsencode(centers_file), int(cosmology), C.byref(res), C.byref(options),
                                C.byref(extra))
    if st != 0:
        raise RuntimeError("countspheres_mocks returned %d" % st)
    pN = np.array([[res.pN[i][j] for j in range(res.num_pN)] for i in range(res.nbin)], dtype=np.float64)
    out = dict(pN=pN, rmax=res.rmax, nbin=res.nbin, nc=res.nc, api_time=options.c_api_time)
    lib.free_results_countspheres_mocks(C.byref(res))
    return out


def call_wp(lib, boxsize, nthreads, pimax, bins, X, Y, Z, w=None, weight_type=None, options=None, dtype=None):
    _declare(lib)
    dtype = np.dtype(dtype or np.asarray(X).dtype)
    X, Y, Z = _prep((X, Y, Z), dtype)
    extra, keep = make_extra(w, w, weight_type, dtype)
    res = ResultsWp()
    with binfile_for(bins) as bf:
        st = lib.countpairs_wp(X.size, _ptr(X), _ptr(Y), _ptr(Z), float(boxsize), int(nthreads), bf, float(pimax),
                               C.byref(res), C.byref(options), C.byref(extra))
    if st != 0:
        raise RuntimeError("countpairs_wp returned %d" % st)
    n = res.nbin
    out = dict(npairs=_arr(res.npairs, n, np.uint64)[1:], rupp=_arr(res.rupp, n, np.float64),
               ravg=_arr(res.rpavg, n, np.float64)[1:], weightavg=_arr(res.weightavg, n, np.float64)[1:],
               cf=_arr(res.wp, n, np.float64)[1:], api_time=options.c_api_time)
    lib.free_results_wp(C.byref(res))
    return out


def call_xi(lib, boxsize, nthreads, bins, X, Y, Z, w=None, weight_type=None, options=None, dtype=None):
    _declare(lib)
    dtype = np.dtype(dtype or np.asarray(X).dtype)
    X, Y, Z = _prep((X, Y, Z), dtype)
    extra, keep = make_extra(w, w, weight_type, dtype)
    res = ResultsXi()
    with binfile_for(bins) as bf:
        st = lib.countpairs_xi(X.size, _ptr(X), _ptr(Y), _ptr(Z), float(boxsize), int(nthreads), bf, C.byref(res),
                               C.byref(options), C.byref(extra))
    if st != 0:
        raise RuntimeError("countpairs_xi returned %d" % st)
    n = res.nbin
    out = dict(npairs=_arr(res.npairs, n, np.uint64)[1:], rupp=_arr(res.rupp, n, np.float64),
               ravg=_arr(res.ravg, n, np.float64)[1:], weightavg=_arr(res.weightavg, n, np.float64)[1:],
               cf=_arr(res.xi, n, np.float64)[1:], api_time=options.c_api_time)
    lib.free_results_xi(C.byref(res))
    return out


def call_DDtheta(lib, autocorr, nthreads, bins, RA1, DEC1, w1=None, RA2=None, DEC2=None, w2=None,
                 weight_type=None, options=None, dtype=None):
    _declare(lib)
    dtype = np.dtype(dtype or np.asarray(RA1).dtype)
    # the C routine may shift RA/DEC in place -> always hand it private copies
    RA1, DEC1, RA2, DEC2 = (None if a is None else np.array(a, dtype=dtype, order="C", copy=True)
                            for a in (RA1, DEC1, RA2, DEC2))
    if autocorr:
        RA2, DEC2 = RA1, DEC1
    extra, keep = make_extra(w1, w2 if not autocorr else w1, weight_type, dtype)
    res = ResultsTheta()
    n2 = 0 if RA2 is None else RA2.size
    with binfile_for(bins) as bf:
        st = lib.countpairs_theta_mocks(RA1.size, _ptr(RA1), _ptr(DEC1), n2, _ptr(RA2), _ptr(DEC2), int(nthreads),
                                        int(autocorr), bf, C.byref(res), C.byref(options), C.byref(extra))
    if st != 0:
        raise RuntimeError("countpairs_theta_mocks returned %d" % st)
    n = res.nbin
    out = dict(npairs=_arr(res.npairs, n, np.uint64)[1:], rupp=_arr(res.theta_upp, n, np.float64),
               ravg=_arr(res.theta_avg, n, np.float64)[1:], weightavg=_arr(res.weightavg, n, np.float64)[1:],
               api_time=options.c_api_time)
    lib.free_results_countpairs_theta(C.byref(res))
    return out
