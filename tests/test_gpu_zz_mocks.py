"""GPU parity tests of the survey-geometry statistics: countpairs_mocks (DDrppi_mocks) and
countpairs_mocks_s_mu (DDsmu_mocks) through the C ABI / the Python drop-in wrappers, against the CPU oracle and the
committed outputs of the unmodified reference (tests/golden/ref_mocks_*.npz).  Bar as in test_gpu_parity.py: npairs
bit-exact, averages within 1e-10 (double) / 1e-5 (float) relative of the oracle's double sums.
(The file name sorts after test_gpu_parity.py: the BASELINE configs are checked first.)"""
import ctypes as C
import os

import numpy as np
import pytest

import harness as H

pytestmark = pytest.mark.gpu

TOL = {np.float64: 1e-10, np.float32: 1e-5}
EDGES = np.logspace(np.log10(0.5), np.log10(30.0), 13)


def _close(a, b, tol, what):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    rel = np.abs(a - b) / np.maximum(np.abs(b), 1e-300)
    rel[(a == 0) & (b == 0)] = 0
    assert rel.max() <= tol, "%s: max rel diff %.3e > %.1e" % (what, rel.max(), tol)


def _gpu(stat, autocorr, dtype, sets, *, pimax=40.0, mu_max=0.9, nmu=10, weights=True, avg=True, edges=EDGES, **okw):
    from corrfunc_b200 import _capi, _lib

    (ra, dec, d, w), (ra2, dec2, d2, w2) = sets
    o = _capi.default_options(dtype, need_avg_sep=avg, is_comoving_dist=True, **okw)
    kw = dict(options=o)
    if weights:
        kw.update(w1=w, weight_type="pair_product")
    if not autocorr:
        kw.update(RA2=ra2, DEC2=dec2, CZ2=d2)
        if weights:
            kw.update(w2=w2)
    if stat == "DDrppi_mocks":
        return _capi.call_DDrppi_mocks(_lib.load(), autocorr, 1, 1, pimax, edges, ra, dec, d, **kw)
    return _capi.call_DDsmu_mocks(_lib.load(), autocorr, 1, 1, mu_max, nmu, edges, ra, dec, d, **kw)


def _oracle(stat, autocorr, sets, *, pimax=40.0, mu_max=0.9, nmu=10, weights=True, avg=True, edges=EDGES, **okw):
    (ra, dec, d, w), (ra2, dec2, d2, w2) = sets
    kw = dict(need_avg=avg, autocorr=bool(autocorr), periodic=False, **okw)
    if weights:
        kw.update(w1=w, weight_type="pair_product")
    if not autocorr:
        kw.update(X2=ra2, Y2=dec2, Z2=d2)
        if weights:
            kw.update(w2=w2)
    if stat == "DDrppi_mocks":
        return H.oracle_theory(stat, ra, dec, d, edges, pimax=pimax, **kw)
    return H.oracle_theory(stat, ra, dec, d, edges, mu_max=mu_max, nmu_bins=nmu, **kw)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("autocorr", [1, 0])
@pytest.mark.parametrize("weights", [False, True])
@pytest.mark.parametrize("stat", ["DDrppi_mocks", "DDsmu_mocks"])
def test_mocks_vs_oracle(stat, weights, autocorr, dtype):
    sets = (H.mock_points(3, 30000, dtype), H.mock_points(4, 20000, dtype))
    g = _gpu(stat, autocorr, dtype, sets, weights=weights, avg=weights)
    a = _oracle(stat, autocorr, sets, weights=weights, avg=weights)
    assert np.array_equal(g["npairs"], a["npairs"])
    if weights:
        _close(g["ravg"], a["ravg"], TOL[dtype], "avg separation")
        _close(g["weightavg"], a["weightavg"], TOL[dtype], "weightavg")


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_mocks_against_reference_golden_outputs(dtype):
    """Committed outputs of the UNMODIFIED reference (AVX-512F kernels) on seeded survey wedges."""
    g = np.load(os.path.join(H.GOLDEN, "ref_mocks_%s.npz" % np.dtype(dtype).name))
    sets = (H.mock_points(int(g["seed"]), int(g["N1"]), dtype), H.mock_points(int(g["seed"]) + 1, int(g["N2"]), dtype))
    tol = TOL[dtype] if dtype == np.float64 else 1e-4  # the reference's float path sums in float
    for autocorr in (1, 0):
        tag = "auto" if autocorr else "cross"
        for stat in ("DDrppi_mocks", "DDsmu_mocks"):
            r = _gpu(stat, autocorr, dtype, sets, pimax=float(g["pimax"]), mu_max=float(g["mu_max"]), nmu=int(g["nmu"]),
                     edges=g["edges"])
            assert np.array_equal(r["npairs"], g["%s_%s__npairs" % (stat, tag)]), (stat, tag)
            _close(r["ravg"], g["%s_%s__ravg" % (stat, tag)], tol, stat + " avg")
            _close(r["weightavg"], g["%s_%s__weightavg" % (stat, tag)], tol, stat + " weightavg")


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("stat", ["DDrppi_mocks", "DDsmu_mocks"])
@pytest.mark.parametrize("occ,refine", [(8, (2, 2, 1)), (40, (1, 1, 1)), (8, (3, 2, 2))])
def test_mocks_subdivided_lattice_and_custom_refine(stat, dtype, occ, refine):
    """Device-side fine cells (target occupancy forced low) and user bin-refine factors leave the counts unchanged."""
    from corrfunc_b200 import _lib

    sets = (H.mock_points(7, 40000, dtype), H.mock_points(8, 10, dtype))
    custom = refine != (2, 2, 1)
    a = _oracle(stat, 1, sets, weights=False, avg=False, refine=refine, custom_refine=custom)
    _lib.load().cfb_set_target_occupancy(occ)
    try:
        g = _gpu(stat, 1, dtype, sets, weights=False, avg=False, bin_refine_factors=refine, custom_refine=custom)
    finally:
        _lib.load().cfb_set_target_occupancy(0)
    assert np.array_equal(g["npairs"], a["npairs"])


def test_mocks_python_wrappers_and_errors():
    from corrfunc_b200 import _capi, _lib
    from corrfunc_b200.mocks import DDrppi_mocks, DDsmu_mocks

    dtype = np.float64
    ra, dec, d, w = H.mock_points(5, 20000, dtype)
    ra_in, dec_in = ra - 180.0, dec + 90.0  # RA in [-180,180], DEC in [0,180]: shifted back like the reference does
    ra, dec = ra_in + 180.0, dec_in - 90.0
    sets = ((ra, dec, d, w), H.mock_points(6, 10, dtype))
    edges = np.logspace(0, np.log10(20.0), 6)
    r = DDrppi_mocks(1, 1, 4, 25.0, edges, ra_in, dec_in, d, weights1=w, weight_type="pair_product",
                     is_comoving_dist=True, output_rpavg=True)
    a = _oracle("DDrppi_mocks", 1, sets, pimax=25.0, edges=edges)
    assert r.dtype.names == ("rmin", "rmax", "rpavg", "pimax", "npairs", "weightavg") and r.size == 5 * 25
    assert np.array_equal(r["npairs"], a["npairs"].ravel())
    assert np.array_equal(r["pimax"][:25], np.arange(1, 26, dtype=np.float64))
    _close(r["rpavg"], a["ravg"].ravel(), 1e-10, "rpavg")
    r, t = DDsmu_mocks(1, 2, 4, 0.8, 4, edges, ra, dec, d, is_comoving_dist=True, c_api_timer=True)
    a = _oracle("DDsmu_mocks", 1, sets, mu_max=0.8, nmu=4, edges=edges, weights=False, avg=False)
    assert r.dtype.names == ("smin", "smax", "savg", "mumax", "npairs", "weightavg") and t > 0
    assert np.array_equal(r["npairs"], a["npairs"].ravel())
    assert np.allclose(r["mumax"][:4], [0.2, 0.4, 0.6, 0.8])
    lib = _lib.load()
    with pytest.raises(RuntimeError):  # init_cosmology: only 1 and 2 exist (utils/cosmology_params.c:21-54)
        _capi.call_DDrppi_mocks(lib, 1, 3, 1, 25.0, edges, ra, dec, d, options=_capi.default_options(dtype, is_comoving_dist=True))
    with pytest.raises(RuntimeError):  # cz = 20 km/s is z < 1e-4: below the distance table (the reference's GSL call aborts)
        _capi.call_DDrppi_mocks(lib, 1, 1, 1, 25.0, edges, ra, dec, np.full_like(d, 20.0), options=_capi.default_options(dtype))
    with pytest.raises(RuntimeError):  # rmin = 0 is not accepted by the mocks statistics
        _capi.call_DDsmu_mocks(lib, 1, 1, 1, 0.8, 4, np.array([0.0, 1.0, 5.0]), ra, dec, d,
                               options=_capi.default_options(dtype, is_comoving_dist=True))


def test_mocks_precision_suffixed_entry_points():
    """countpairs_mocks_float / countpairs_mocks_s_mu_double (the *_impl.h.src prototypes): typed pointers,
    options->float_type ignored; pimax arrives as a float in the float variant."""
    from corrfunc_b200 import _capi, _lib

    lib = _lib.load()
    _capi._declare(lib)
    edges = np.logspace(0, np.log10(20.0), 6)
    for dtype, suf in ((np.float32, "float"), (np.float64, "double")):
        ra, dec, d, _ = H.mock_points(9, 15000, dtype)
        sets = ((ra, dec, d, None), (None, None, None, None))
        want = _gpu("DDrppi_mocks", 1, dtype, sets, pimax=30.0, edges=edges, weights=False, avg=False)
        o = _capi.default_options(dtype, is_comoving_dist=True)
        o.float_type = 12 - o.float_type
        e, keep = _capi.make_extra(None, None, None, dtype)
        r = _capi.ResultsMocksRpPi()
        fn = getattr(lib, "countpairs_mocks_" + suf)
        fn.restype = C.c_int
        fn.argtypes = None
        with _capi.binfile_for(edges) as bf:
            p = [C.c_void_p(a.ctypes.data) for a in (ra, dec, d)]
            pim = C.c_float(30.0) if dtype == np.float32 else C.c_double(30.0)
            st = fn(C.c_int64(ra.size), p[0], p[1], p[2], C.c_int64(ra.size), p[0], p[1], p[2], C.c_int(1), C.c_int(1), bf,
                    pim, C.c_int(1), C.byref(r), C.byref(o), C.byref(e))
        assert st == 0
        tot = (r.nbin + 1) * (r.npibin + 1)
        got = np.ctypeslib.as_array(r.npairs, shape=(tot,)).reshape(r.nbin + 1, r.npibin + 1)[1:r.nbin, :r.npibin].copy()
        lib.free_results_mocks(C.byref(r))
        assert np.array_equal(got, want["npairs"])
        assert o.float_type == 12 - np.dtype(dtype).itemsize
