"""The HOST layer of countspheres_mocks without a GPU: cf_host.c (the product's own host code) is linked with
tests/stub_device/stub_device.c -- a brute-force CPU stand-in for the CUDA layer's C ABI, test infrastructure only --
into a throw-away library, and driven through the same ctypes helpers as the real one.  Checks what the CUDA kernel
does not decide: centres file parsing, cz -> distance, the shift, the choice of centres on the randoms and the
rewritten file, cumulative counts and pN.  (The kernel itself is checked by the -m gpu tests.)"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import harness as H
from corrfunc_b200 import _capi


@pytest.fixture(scope="module")
def hostlib(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("hoststub") / "libcorrfunc_hoststub.so")
    host = os.path.join(H.ROOT, "corrfunc_b200", "csrc", "host")
    subprocess.check_call(["/usr/bin/gcc", "-std=c11", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-fopenmp",
                           "-I", os.path.join(H.ROOT, "include"), "-I", host, os.path.join(host, "cf_host.c"),
                           os.path.join(H.ROOT, "tests", "stub_device", "stub_device.c"), "-o", out, "-lm"])
    return C.CDLL(out, mode=os.RTLD_LOCAL)


def test_host_layer_reproduces_the_reference_golden_vpf(hostlib):
    ra, dec, cz, _ = H.load_mr19_mock_cz()
    o = _capi.default_options(np.float64, bin_refine_factors=(1, 1, 1))
    r = _capi.call_vpf_mocks(hostlib, 10.0, 10, 10000, 6, 1, H.VPF_CENTERS, 1, ra, dec, cz, options=o)
    assert np.allclose(r["pN"], H.load_vpf_golden(), atol=1e-9, rtol=1e-6)
    assert r["nbin"] == 10 and r["nc"] == 10000 and r["rmax"] == 10.0


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_host_layer_places_centres_on_the_randoms(hostlib, dtype, tmp_path):
    ra, dec, d, _ = H.mock_points(41, 20000, dtype)
    rra, rdec, rd, _ = H.mock_points(42, 3000, dtype)
    cfile = str(tmp_path / "centres.txt")
    nc = 150
    o = _capi.default_options(dtype, bin_refine_factors=(1, 1, 1), is_comoving_dist=True, c_api_timer=True)
    r = _capi.call_vpf_mocks(hostlib, 12.0, 6, nc, 4, 2, cfile, 1, ra, dec, d, RAND_RA=rra, RAND_DEC=rdec, RAND_CZ=rd,
                             options=o)
    rcube = dtype(max(d.max(), rd.max())) + dtype(1.0)
    xc, yc, zc = H.vpf_centres_from_randoms(rra, rdec, rd, rcube, 12.0, 2, nc)
    written = np.loadtxt(cfile)
    assert written.shape == (nc, 4) and np.allclose(written[:, 0], xc, atol=1e-4) and np.all(written[:, 3] == 12.0)
    want, _ = H.oracle_vpf_mocks(ra, dec, d, xc, yc, zc, 12.0, 6, 4, dmax_randoms=rd.max())
    assert np.array_equal(r["pN"], want)
    assert r["api_time"] > 0
    # errors are loud: unknown cosmology, too few randoms to place a single sphere
    with pytest.raises(RuntimeError):
        _capi.call_vpf_mocks(hostlib, 12.0, 6, nc, 4, 2, cfile, 7, ra, dec, d, RAND_RA=rra, RAND_DEC=rdec, RAND_CZ=rd,
                             options=_capi.default_options(dtype, is_comoving_dist=True))
    with pytest.raises(RuntimeError):
        _capi.call_vpf_mocks(hostlib, 12.0, 6, nc, 4, 10 ** 6, str(tmp_path / "none.txt"), 1, ra, dec, d, RAND_RA=rra,
                             RAND_DEC=rdec, RAND_CZ=rd, options=_capi.default_options(dtype, is_comoving_dist=True))


# ---- host layer of countpairs_mocks / countpairs_mocks_s_mu (angles, cz, extents, bins, epilogue) -----------------

def test_host_layer_reproduces_the_reference_golden_DDrppi_mocks(hostlib):
    """mocks/tests/Mr19_mock.DD (cz input, written upstream with real GSL) through the product's host code."""
    ra, dec, cz, w = H.load_mr19_mock_cz()
    bins = H.load_bins_file("mocks_bins.txt")
    gold = H.load_ddrppi_mocks_golden()
    o = _capi.default_options(np.float64, need_avg_sep=True, is_comoving_dist=False)
    r = _capi.call_DDrppi_mocks(hostlib, 1, 1, 4, 40.0, bins, ra, dec, cz, w1=w, weight_type="pair_product", options=o)
    assert np.array_equal(r["npairs"].ravel(), gold["npairs"])
    assert np.allclose(r["ravg"].ravel(), gold["ravg"], atol=1e-9, rtol=1e-6)
    assert np.allclose(r["weightavg"].ravel(), gold["weightavg"], atol=1e-9, rtol=1e-6)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_host_layer_mocks_vs_committed_reference_outputs(hostlib, dtype):
    g = np.load(os.path.join(H.GOLDEN, "ref_mocks_%s.npz" % np.dtype(dtype).name))
    n1, n2 = 20000, 12000  # a subsample keeps the brute force short: compared with the oracle, not the goldens
    ra, dec, d, w = H.mock_points(int(g["seed"]), n1, dtype)
    ra2, dec2, d2, w2 = H.mock_points(int(g["seed"]) + 1, n2, dtype)
    cz, cz2 = (d * dtype(60.0)).astype(dtype), (d2 * dtype(60.0)).astype(dtype)
    o = _capi.default_options(dtype, need_avg_sep=True, is_comoving_dist=False)
    r = _capi.call_DDsmu_mocks(hostlib, 0, 2, 1, 0.9, 10, g["edges"], ra - dtype(180.0), dec + dtype(90.0), cz, w1=w,
                               RA2=ra2, DEC2=dec2, CZ2=cz2, w2=w2, weight_type="pair_product", options=o)
    raf, decf = (ra - dtype(180.0)) + dtype(180.0), (dec + dtype(90.0)) - dtype(90.0)  # shifted back in place by the host layer
    czmax = max(cz.max(), cz2.max())
    D = H.cz_to_comoving(np.append(cz, czmax).astype(dtype), 2)[:-1]
    D2 = H.cz_to_comoving(np.append(cz2, czmax).astype(dtype), 2)[:-1]
    a = H.oracle_theory("DDsmu_mocks", raf, decf, D, g["edges"], mu_max=0.9, nmu_bins=10, autocorr=False, X2=ra2, Y2=dec2,
                        Z2=D2, w1=w, w2=w2, weight_type="pair_product", need_avg=True, periodic=False)
    assert np.array_equal(r["npairs"], a["npairs"])
    tol = 1e-10 if dtype == np.float64 else 1e-5
    ok = a["npairs"] > 0
    assert np.allclose(r["ravg"][ok], a["ravg"][ok], rtol=tol) and np.allclose(r["weightavg"][ok], a["weightavg"][ok], rtol=tol)


# ---- theory vpf: MT19937 stream, centres, host driver -------------------------------------------------------------

def test_mt19937_stream_is_the_published_one():
    """corrfunc_b200_mt19937_uniform (what countspheres draws its centres from) against numpy's independent MT19937
    with the same (2002) initialisation, which is the stream gsl_rng_mt19937 + gsl_rng_uniform produce."""
    for seed in (42, 1, 0):
        bg = np.random.MT19937()
        bg._legacy_seeding(seed if seed else 4357)  # GSL maps seed 0 to 4357
        want = bg.random_raw(3000) / 4294967296.0
        assert np.array_equal(H.mt19937_uniform(seed, 3000), want)


@pytest.mark.skipif(H.load_ref() is None or not hasattr(H.load_ref(), "countspheres"),
                    reason="oracle/_ref was not prebuilt with theory/vpf")
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("periodic", [True, False])
def test_host_layer_theory_vpf_vs_live_reference(hostlib, dtype, periodic):
    """theory/vpf of the unmodified reference (its GSL generator supplied by oracle/gsl_shim) vs the product's host
    layer over the brute-force device stand-in vs the oracle on the centres of tests/harness.py."""
    L, N = 300.0, 40000
    x, y, z, _ = H.box_points(9, N, L, dtype)
    rmax, nbin, nc, num_pN, seed = 12.0, 6, 800, 5, 77
    o = _capi.default_options(dtype, periodic=periodic, boxsize=L if periodic else None, isa=H.ref_isa(), bin_refine_factors=(1, 1, 1))
    r = _capi.call_vpf(H.load_ref(), rmax, nbin, nc, num_pN, seed, x, y, z, options=o)
    o = _capi.default_options(dtype, periodic=periodic, boxsize=L if periodic else None, bin_refine_factors=(1, 1, 1))
    h = _capi.call_vpf(hostlib, rmax, nbin, nc, num_pN, seed, x, y, z, options=o)
    assert np.array_equal(h["pN"], r["pN"])
    xc, yc, zc, wrap = H.vpf_theory_centres(x, y, z, rmax, nc, seed, periodic, L)
    a = H.oracle_vpf_theory(x, y, z, xc, yc, zc, periodic, wrap, rmax, nbin, num_pN)
    assert np.array_equal(a, r["pN"])
    assert abs(r["pN"][:, :].sum(axis=1).max() - 1.0) < 0.5 and r["pN"][0, 0] > r["pN"][-1, 0]  # sanity: p0 falls with radius


@pytest.mark.skipif(H.load_ref() is None or not hasattr(H.load_ref(), "countspheres"),
                    reason="oracle/_ref was not prebuilt with theory/vpf")
def test_theory_vpf_small_periodic_box_vs_live_reference(hostlib):
    """Three cells per axis, spheres reaching through the periodic faces: nearest-image brute force == the reference's
    per-cell centre shifts."""
    x, y, z, _ = H.box_points(10, 3000, 40.0, np.float64)
    o = _capi.default_options(np.float64, periodic=True, boxsize=40.0, isa=H.ref_isa(), bin_refine_factors=(1, 1, 1))
    r = _capi.call_vpf(H.load_ref(), 12.0, 4, 300, 4, 5, x, y, z, options=o)
    xc, yc, zc, wrap = H.vpf_theory_centres(x, y, z, 12.0, 300, 5, True, 40.0)
    assert np.array_equal(H.oracle_vpf_theory(x, y, z, xc, yc, zc, True, wrap, 12.0, 4, 4), r["pN"])
    o = _capi.default_options(np.float64, periodic=True, boxsize=40.0, bin_refine_factors=(1, 1, 1))
    assert np.array_equal(_capi.call_vpf(hostlib, 12.0, 4, 300, 4, 5, x, y, z, options=o)["pN"], r["pN"])


def test_host_layer_mocks_error_behaviour(hostlib):
    """What the host layer decides before any device work: the reference's error returns, loudly."""
    ra, dec, d, _ = H.mock_points(5, 2000, np.float64)
    edges = np.logspace(0, 1.3, 6)
    ok = lambda **kw: _capi.default_options(np.float64, **kw)  # noqa: E731
    with pytest.raises(RuntimeError):  # init_cosmology knows 1 and 2 only
        _capi.call_DDrppi_mocks(hostlib, 1, 3, 1, 25.0, edges, ra, dec, d, options=ok(is_comoving_dist=True))
    with pytest.raises(RuntimeError):  # cz = 20 km/s is z < 1e-4: below the distance table
        _capi.call_DDrppi_mocks(hostlib, 1, 1, 1, 25.0, edges, ra, dec, np.full_like(d, 20.0), options=ok())
    with pytest.raises(RuntimeError):  # rmin = 0 is refused by the mocks statistics
        _capi.call_DDsmu_mocks(hostlib, 1, 1, 1, 0.8, 4, np.array([0.0, 1.0, 5.0]), ra, dec, d, options=ok(is_comoving_dist=True))
    with pytest.raises(RuntimeError):  # mu_max outside (0, 1]
        _capi.call_DDsmu_mocks(hostlib, 1, 1, 1, 1.5, 4, edges, ra, dec, d, options=ok(is_comoving_dist=True))
    with pytest.raises(RuntimeError):  # DEC beyond 180 degrees
        _capi.call_DDrppi_mocks(hostlib, 1, 1, 1, 25.0, edges, ra, dec + 200.0, d, options=ok(is_comoving_dist=True))
    with pytest.raises(RuntimeError):  # pimax below one bin
        _capi.call_DDrppi_mocks(hostlib, 1, 1, 1, 0.5, edges, ra, dec, d, options=ok(is_comoving_dist=True))
    empty = np.zeros(0)
    r = _capi.call_DDrppi_mocks(hostlib, 1, 1, 1, 25.0, edges, empty, empty, empty, options=ok(is_comoving_dist=True))
    assert r["npairs"].size == 0  # the reference returns EXIT_SUCCESS and leaves the results untouched
    with pytest.raises(RuntimeError):  # vpf: nonsense parameters
        _capi.call_vpf(hostlib, -1.0, 4, 10, 3, 1, ra, dec, d, options=ok(periodic=False))


def test_host_layer_interrupt(hostlib, capfd, monkeypatch):
    """utils/macros.h:145-167, theory/DD/countpairs_impl.c.src:31-37,554-569: a SIGINT / SIGTERM / SIGHUP that arrives during
    a call is caught by the library's own handler (message on stderr, flag for the device), the call returns
    EXIT_FAILURE, and the caller's handlers are back in place afterwards."""
    import signal

    ra, dec, d, _ = H.mock_points(5, 2000, np.float64)
    bins = np.logspace(-1, 1, 6)
    o = _capi.default_options(np.float64, is_comoving_dist=True)
    ok = _capi.call_DDrppi_mocks(hostlib, 1, 1, 1, 10.0, bins, ra, dec, d, options=o)
    assert ok["npairs"].sum() > 0
    for signo in (signal.SIGINT, signal.SIGTERM, signal.SIGHUP):
        monkeypatch.setenv("CFB_STUB_RAISE", str(int(signo)))
        with pytest.raises(RuntimeError):
            _capi.call_DDrppi_mocks(hostlib, 1, 1, 1, 10.0, bins, ra, dec, d, options=_capi.default_options(np.float64, is_comoving_dist=True))
        err = capfd.readouterr().err
        assert "Received signal" in err and "Aborting" in err and "signo = %d" % int(signo) in err
    monkeypatch.delenv("CFB_STUB_RAISE")
    # the flag is cleared and the handlers are restored: the next call succeeds, and Python sees its own SIGINT again
    again = _capi.call_DDrppi_mocks(hostlib, 1, 1, 1, 10.0, bins, ra, dec, d, options=_capi.default_options(np.float64, is_comoving_dist=True))
    assert np.array_equal(again["npairs"], ok["npairs"])
    with pytest.raises(KeyboardInterrupt):
        signal.raise_signal(signal.SIGINT)
        import time
        time.sleep(0.01)
