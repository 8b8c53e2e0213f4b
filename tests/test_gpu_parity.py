"""GPU parity tests: the CUDA path, called through the C ABI / the Python drop-in wrappers, against
  (1) the CPU oracle (oracle/libpairs_oracle.so) on the same seeded inputs,
  (2) committed golden outputs of the unmodified reference (tests/golden/ref_synthetic_*.npz, the
      reference's own Mr19 DDtheta golden file),
  (3) the unmodified reference itself (oracle/_ref) when the prebuilt library travelled to the box,
  (4) the reference's data-free known-answer tests (Corrfunc/tests/test_theory.py:115-286).
Bar: npairs bit-exact (double AND float on these inputs); averages within 1e-10 (double) / 1e-5 (float)
relative -- the tolerances BASELINE.json's north_star states.
"""
import os

import numpy as np
import pytest

import harness as H

pytestmark = pytest.mark.gpu

TOL = {np.float64: 1e-10, np.float32: 1e-5}


def _theory():
    import corrfunc_b200.theory as T

    return T


def _close(a, b, tol, what):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    scale = np.maximum(np.abs(b), 1e-300)
    rel = np.abs(a - b) / scale
    # bins with no pairs hold 0 in both
    rel[(a == 0) & (b == 0)] = 0
    assert rel.max() <= tol, "%s: max rel diff %.3e > %.1e" % (what, rel.max(), tol)


@pytest.fixture(scope="module")
def edges():
    return H.load_bins_file("theory_bins.txt")


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("periodic", [True, False])
@pytest.mark.parametrize("autocorr", [1, 0])
@pytest.mark.parametrize("weights", [False, True])
def test_DD_vs_oracle(dtype, periodic, autocorr, weights, edges):
    T = _theory()
    L, N = 420.0, 30000
    x, y, z, w = H.box_points(11, N, L, dtype)
    x2, y2, z2, w2 = H.box_points(12, N // 2, L, dtype)
    kw = dict(periodic=periodic, boxsize=L, output_ravg=True)
    okw = dict(periodic=periodic, boxsize=L, need_avg=True, autocorr=bool(autocorr))
    if weights:
        kw.update(weights1=w, weight_type="pair_product")
        okw.update(w1=w, weight_type="pair_product")
    if not autocorr:
        kw.update(X2=x2, Y2=y2, Z2=z2)
        okw.update(X2=x2, Y2=y2, Z2=z2)
        if weights:
            kw.update(weights2=w2)
            okw.update(w2=w2)
    got = T.DD(autocorr, 4, edges, x, y, z, **kw)
    ref = H.oracle_theory("DD", x, y, z, edges, **okw)
    assert np.array_equal(got["npairs"], ref["npairs"])
    _close(got["ravg"], ref["ravg"], TOL[dtype], "ravg")
    if weights:
        _close(got["weightavg"], ref["weightavg"], TOL[dtype], "weightavg")


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("periodic", [True, False])
@pytest.mark.parametrize("autocorr", [1, 0])
def test_DDrppi_vs_oracle(dtype, periodic, autocorr, edges):
    T = _theory()
    L, N, pimax = 420.0, 30000, 40.0
    x, y, z, w = H.box_points(21, N, L, dtype)
    x2, y2, z2, w2 = H.box_points(22, N // 2, L, dtype)
    kw = dict(periodic=periodic, boxsize=L, output_rpavg=True, weights1=w, weight_type="pair_product")
    okw = dict(periodic=periodic, boxsize=L, need_avg=True, autocorr=bool(autocorr), w1=w,
               weight_type="pair_product", pimax=pimax)
    if not autocorr:
        kw.update(X2=x2, Y2=y2, Z2=z2, weights2=w2)
        okw.update(X2=x2, Y2=y2, Z2=z2, w2=w2)
    got = T.DDrppi(autocorr, 4, pimax, edges, x, y, z, **kw)
    ref = H.oracle_theory("DDrppi", x, y, z, edges, **okw)
    assert np.array_equal(got["npairs"], ref["npairs"].ravel())
    _close(got["rpavg"], ref["ravg"].ravel(), TOL[dtype], "rpavg")
    _close(got["weightavg"], ref["weightavg"].ravel(), TOL[dtype], "weightavg")


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("periodic", [True, False])
@pytest.mark.parametrize("autocorr", [1, 0])
def test_DDsmu_vs_oracle(dtype, periodic, autocorr, edges):
    T = _theory()
    L, N, mu_max, nmu = 420.0, 30000, 0.5, 10
    x, y, z, w = H.box_points(31, N, L, dtype)
    x2, y2, z2, w2 = H.box_points(32, N // 2, L, dtype)
    kw = dict(periodic=periodic, boxsize=L, output_savg=True, weights1=w, weight_type="pair_product")
    okw = dict(periodic=periodic, boxsize=L, need_avg=True, autocorr=bool(autocorr), w1=w,
               weight_type="pair_product", mu_max=mu_max, nmu_bins=nmu)
    if not autocorr:
        kw.update(X2=x2, Y2=y2, Z2=z2, weights2=w2)
        okw.update(X2=x2, Y2=y2, Z2=z2, w2=w2)
    got = T.DDsmu(autocorr, 4, edges, mu_max, nmu, x, y, z, **kw)
    ref = H.oracle_theory("DDsmu", x, y, z, edges, **okw)
    assert np.array_equal(got["npairs"], ref["npairs"].ravel())
    _close(got["savg"], ref["ravg"].ravel(), TOL[dtype], "savg")
    _close(got["weightavg"], ref["weightavg"].ravel(), TOL[dtype], "weightavg")


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("weights", [False, True])
def test_xi_wp_vs_oracle(dtype, weights, edges):
    T = _theory()
    L, N, pimax = 420.0, 40000, 40.0
    x, y, z, w = H.box_points(41, N, L, dtype)
    kw = dict(weights=w, weight_type="pair_product") if weights else {}
    okw = dict(w1=w, weight_type="pair_product") if weights else {}
    got = T.xi(L, 4, edges, x, y, z, output_ravg=True, **kw)
    ref = H.oracle_theory("xi", x, y, z, edges, boxsize=L, need_avg=True, **okw)
    assert np.array_equal(got["npairs"], ref["npairs"])
    _close(got["ravg"], ref["ravg"], TOL[dtype], "ravg")
    assert np.allclose(got["xi"], ref["cf"], rtol=1e-6 if dtype == np.float64 else 1e-2, atol=1e-9 if dtype == np.float64 else 1e-3)
    got = T.wp(L, pimax, 4, edges, x, y, z, output_rpavg=True, **kw)
    ref = H.oracle_theory("wp", x, y, z, edges, boxsize=L, pimax=pimax, need_avg=True, **okw)
    assert np.array_equal(got["npairs"], ref["npairs"])
    _close(got["rpavg"], ref["ravg"], TOL[dtype], "rpavg")
    assert np.allclose(got["wp"], ref["cf"], rtol=1e-6 if dtype == np.float64 else 1e-2, atol=1e-7 if dtype == np.float64 else 1e-1)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_subdivided_lattice_matches(dtype, edges):
    """Dense cells force the GPU's fine (sub-divided) lattice; results must not change."""
    from corrfunc_b200 import _lib

    T = _theory()
    L, N = 100.0, 60000
    x, y, z, w = H.box_points(51, N, L, dtype)
    bins = np.logspace(-1, np.log10(20.0), 12)
    ref = H.oracle_theory("DD", x, y, z, bins, periodic=True, boxsize=L, need_avg=True, w1=w, weight_type="pair_product")
    lib = _lib.load()
    for occ in (16, 48, 0):
        lib.cfb_set_target_occupancy(occ)
        got = T.DD(1, 4, bins, x, y, z, periodic=True, boxsize=L, output_ravg=True, weights1=w, weight_type="pair_product")
        st = _lib.last_stats()
        assert np.array_equal(got["npairs"], ref["npairs"]), (occ, st)
        _close(got["ravg"], ref["ravg"], TOL[dtype], "ravg")
    lib.cfb_set_target_occupancy(0)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_against_reference_golden_outputs(dtype):
    """Committed outputs of the UNMODIFIED reference (AVX-512 kernels) on seeded synthetic inputs."""
    T = _theory()
    g = np.load(os.path.join(H.GOLDEN, "ref_synthetic_%s.npz" % np.dtype(dtype).name))
    seed, N, L, edges = int(g["seed"]), int(g["N"]), float(g["L"]), g["edges"]
    x, y, z, w = H.box_points(seed, N, L, dtype)
    x2, y2, z2, w2 = H.box_points(seed + 1, N // 2, L, dtype)
    tol = TOL[dtype] if dtype == np.float64 else 1e-4  # the reference's float path sums in float
    for periodic in (True, False):
        p = "per" if periodic else "nonper"
        r = T.DD(1, 4, edges, x, y, z, weights1=w, weight_type="pair_product", periodic=periodic, boxsize=L, output_ravg=True)
        assert np.array_equal(r["npairs"], g["DD_auto_%s__npairs" % p])
        _close(r["ravg"], g["DD_auto_%s__ravg" % p], tol, "ravg")
        _close(r["weightavg"], g["DD_auto_%s__weightavg" % p], tol, "weightavg")
        r = T.DD(0, 4, edges, x, y, z, weights1=w, weight_type="pair_product", periodic=periodic, boxsize=L,
                 output_ravg=True, X2=x2, Y2=y2, Z2=z2, weights2=w2)
        assert np.array_equal(r["npairs"], g["DD_cross_%s__npairs" % p])
        r = T.DDrppi(1, 4, 40.0, edges, x, y, z, weights1=w, weight_type="pair_product", periodic=periodic, boxsize=L, output_rpavg=True)
        assert np.array_equal(r["npairs"], g["DDrppi_auto_%s__npairs" % p].ravel())
        _close(r["rpavg"], g["DDrppi_auto_%s__ravg" % p].ravel(), tol, "rpavg")
        r = T.DDsmu(1, 4, edges, 0.5, 10, x, y, z, weights1=w, weight_type="pair_product", periodic=periodic, boxsize=L, output_savg=True)
        assert np.array_equal(r["npairs"], g["DDsmu_auto_%s__npairs" % p].ravel())
        _close(r["savg"], g["DDsmu_auto_%s__ravg" % p].ravel(), tol, "savg")
    r = T.xi(L, 4, edges, x, y, z, weights=w, weight_type="pair_product", output_ravg=True)
    assert np.array_equal(r["npairs"], g["xi__npairs"])
    r = T.wp(L, 40.0, 4, edges, x, y, z, weights=w, weight_type="pair_product", output_rpavg=True)
    assert np.array_equal(r["npairs"], g["wp__npairs"])


def test_DDtheta_reference_golden_file():
    """The reference's own known-answer test (Corrfunc/tests/test_mocks.py:61-80): DDtheta autocorr of
    the Mr19 mock vs mocks/tests/Mr19_mock_wtheta.DD, atol 1e-9 / rtol 1e-6 as in common.py:83-105."""
    from corrfunc_b200.mocks import DDtheta_mocks

    ra, dec, w = H.load_mr19_mock()
    bins = H.load_bins_file("angular_bins.txt")
    gold = H.load_wtheta_golden()
    r = DDtheta_mocks(1, 4, bins, ra, dec, weights1=w, weight_type="pair_product", output_thetaavg=True)
    assert np.array_equal(r["npairs"], gold["npairs"])
    assert np.allclose(r["thetaavg"], gold["ravg"], atol=1e-9, rtol=1e-6)
    assert np.allclose(r["weightavg"], gold["weightavg"], atol=1e-9, rtol=1e-6)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("autocorr", [1, 0])
@pytest.mark.parametrize("link", [(1, 1), (1, 0), (0, 0)])
def test_DDtheta_vs_oracle(dtype, autocorr, link):
    from corrfunc_b200.mocks import DDtheta_mocks

    ra1, dec1 = H.sphere_points(5, 30000, dtype)
    ra2, dec2 = H.sphere_points(6, 20000, dtype)
    w1 = (1.0 - np.random.default_rng(7).random(ra1.size)).astype(dtype)
    w2 = (1.0 - np.random.default_rng(8).random(ra2.size)).astype(dtype)
    # theta_min large enough that cos(theta_min) < 1 in float (the reference warns below 0.2 deg)
    tb = np.logspace(np.log10(0.05), 1, 16)
    kw = dict(weights1=w1, weight_type="pair_product", output_thetaavg=True, link_in_dec=bool(link[0]), link_in_ra=bool(link[1]))
    okw = dict(w1=w1, weight_type="pair_product", need_avg=True, link_in_dec=link[0], link_in_ra=link[1], autocorr=bool(autocorr))
    if not autocorr:
        kw.update(RA2=ra2, DEC2=dec2, weights2=w2)
        okw.update(RA2=ra2, DEC2=dec2, w2=w2)
    got = DDtheta_mocks(autocorr, 4, tb, ra1, dec1, **kw)
    ref = H.oracle_theta(ra1, dec1, tb, **okw)
    assert np.array_equal(got["npairs"], ref["npairs"])
    _close(got["thetaavg"], ref["ravg"], 1e-9 if dtype == np.float64 else 1e-4, "thetaavg")
    _close(got["weightavg"], ref["weightavg"], TOL[dtype], "weightavg")


# ---- the reference's data-free known-answer tests -------------------------------------------------

@pytest.mark.parametrize("N", [1, 2])
def test_narrow_extent(N):
    T = _theory()
    boxsize = (3.0, 3.0, 3.0)
    r_bins = [0.2, 0.6, 1.0]
    pos = np.array([[0.0, 0.0], [0.0, 0.0], [0.0, 0.5]]) if N == 2 else np.array([[0.1], [0.2], [0.3]])
    res = T.DD(1, 1, r_bins, pos[0], pos[1], pos[2], boxsize=boxsize, periodic=True)
    assert np.all(res["npairs"] == ([2, 0] if N == 2 else [0, 0]))


@pytest.mark.parametrize("autocorr", [0, 1])
@pytest.mark.parametrize("binref", [1, 2, 3])
@pytest.mark.parametrize("min_sep_opt", [False, True])
@pytest.mark.parametrize("maxcells", [1, 2, 3])
def test_duplicate_cellpairs(autocorr, binref, min_sep_opt, maxcells):
    T = _theory()
    boxsize = 432.0
    kw = dict(boxsize=boxsize, periodic=True, max_cells_per_dim=maxcells, xbin_refine_factor=binref,
              ybin_refine_factor=binref, zbin_refine_factor=binref, enable_min_sep_opt=min_sep_opt)
    r_bins = np.array([0.01, 0.4]) * boxsize
    pos = np.array([[0.02, 0.98], [0.0, 0.0], [0.0, 0.0]]) * boxsize
    res = T.DD(autocorr, 1, r_bins, pos[0], pos[1], pos[2], X2=pos[0], Y2=pos[1], Z2=pos[2], **kw)
    assert np.all(res["npairs"] == [2])
    r_bins = np.array([0.2, 0.3, 0.49]) * boxsize
    pos = np.array([[0.0, 0.0], [0.0, 0.0], [0.0, 0.48]]) * boxsize
    res = T.DD(autocorr, 1, r_bins, pos[0], pos[1], pos[2], X2=pos[0], Y2=pos[1], Z2=pos[2], **kw)
    assert np.all(res["npairs"] == [0, 2])


@pytest.mark.parametrize("autocorr", [0, 1], ids=["cross", "auto"])
@pytest.mark.parametrize("binref", [1, 2, 3], ids=["ref1", "ref2", "ref3"])
@pytest.mark.parametrize("maxcells", [1, 2, 3], ids=["max1", "max2", "max3"])
@pytest.mark.parametrize("boxsize", [123.0, (51.0, 75.0, 123.0)], ids=["iso", "aniso"])
@pytest.mark.parametrize("funcname", ["DD", "DDrppi", "DDsmu"])
@pytest.mark.parametrize("periodic", [False, True], ids=["nowrap", "wrap"])
def test_brute(autocorr, binref, maxcells, boxsize, funcname, periodic):
    """Corrfunc/tests/test_theory.py:197-286: two small clouds, numpy brute-force histogram."""
    T = _theory()
    np.random.seed(1234)
    npts, eps = 100, 0.2
    boxsize = np.array(boxsize)
    bins = np.linspace(0.01, 0.49 * boxsize.min(), 20) if periodic else np.linspace(0.01, 2 * boxsize.max(), 20)
    pimax = np.floor(0.49 * boxsize.min())
    mu_max, nmu_bins = 0.5, 10
    func = getattr(T, funcname)
    pos = np.random.uniform(low=-eps, high=eps, size=(npts, 3)) * boxsize
    pos[npts // 2:] += boxsize / 2.0
    pos %= boxsize
    pdiff = np.abs(pos[:, np.newaxis] - pos)
    if periodic:
        mask = pdiff >= boxsize / 2
        pdiff -= mask * boxsize
    args = [autocorr, 1, bins, pos[:, 0], pos[:, 1], pos[:, 2]]
    kwargs = dict(periodic=periodic, boxsize=boxsize, X2=pos[:, 0], Y2=pos[:, 1], Z2=pos[:, 2],
                  max_cells_per_dim=maxcells, xbin_refine_factor=binref, ybin_refine_factor=binref,
                  zbin_refine_factor=binref)
    if funcname == "DDrppi":
        args.insert(2, pimax)
        sqr_rp = (pdiff[:, :, :2] ** 2).sum(axis=-1).reshape(-1)
        pidiff = np.abs(pdiff[:, :, 2]).reshape(-1)
        pibins = np.linspace(0.0, pimax, int(pimax) + 1)
        brute, _, _ = np.histogram2d(sqr_rp, pidiff, bins=(bins ** 2, pibins))
        brute = brute.reshape(-1)
    elif funcname == "DDsmu":
        args[3:3] = (mu_max, nmu_bins)
        sdiff = np.sqrt((pdiff ** 2).sum(axis=-1).reshape(-1))
        sdiff[sdiff == 0.0] = np.inf
        mu = np.abs(pdiff[:, :, 2]).reshape(-1) / sdiff
        mubins = np.linspace(0, mu_max, nmu_bins + 1)
        brute, _, _ = np.histogram2d(sdiff, mu, bins=(bins, mubins))
        brute = brute.reshape(-1)
    else:
        brute, _ = np.histogram((pdiff ** 2).sum(axis=-1).reshape(-1), bins=bins ** 2)
    assert np.any(brute > 0)
    res = func(*args, **kwargs)
    assert np.all(res["npairs"] == brute)


def test_errors_are_loud():
    T = _theory()
    x = np.random.default_rng(0).random(1000) * 10.0
    with pytest.raises(RuntimeError):  # rmax >= L/2 under periodic wrap (gridlink_utils.c.src:37-41)
        T.DD(1, 1, np.linspace(0.1, 6.0, 5), x, x, x, boxsize=10.0, periodic=True)
    with pytest.raises(RuntimeError):  # particles outside [0, L] for xi (gridlink_impl.c.src:183)
        T.xi(5.0, 1, np.linspace(0.1, 1.0, 5), x, x, x)


def test_full_size_config1_properties():
    """BASELINE config 1 at full size (1.2M points, L=420, 14 log bins 0.1-25, double, autocorr):
    size-independent checks -- DD == xi counts (different lattice extents, same pairs), cross(D,D)
    equals auto(D) (no self pairs since rmin>0), and the total is invariant under the device lattice."""
    from corrfunc_b200 import _lib

    T = _theory()
    N, L = 1200000, 420.0
    x, y, z, _ = H.box_points(1001, N, L, np.float64)
    bins = np.logspace(np.log10(0.1), np.log10(25.0), 15)
    a = T.DD(1, 4, bins, x, y, z, periodic=True, boxsize=L)
    b = T.xi(L, 4, bins, x, y, z)
    assert np.array_equal(a["npairs"], b["npairs"])
    c = T.DD(0, 4, bins, x, y, z, X2=x, Y2=y, Z2=z, periodic=True, boxsize=L)
    assert np.array_equal(a["npairs"], c["npairs"])
    _lib.load().cfb_set_target_occupancy(24)
    d = T.DD(1, 4, bins, x, y, z, periodic=True, boxsize=L)
    _lib.load().cfb_set_target_occupancy(0)
    assert np.array_equal(a["npairs"], d["npairs"])
    # expected pair count for a uniform periodic box: N(N-1) * V_shell / L^3 within a few sigma
    vol = 4.0 / 3.0 * np.pi * (bins[1:] ** 3 - bins[:-1] ** 3)
    expect = N * (N - 1.0) * vol / L ** 3
    assert np.all(np.abs(a["npairs"] - expect) < 6 * np.sqrt(expect) + 10)


@pytest.mark.skipif(H.load_ref() is None, reason="oracle/_ref was not prebuilt")
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_live_reference(dtype, edges):
    """The unmodified reference, run on this box's CPU, vs the GPU on a 200k-point box."""
    from corrfunc_b200 import _capi as capi

    T = _theory()
    ref = H.load_ref()
    L, N = 420.0, 200000
    x, y, z, w = H.box_points(77, N, L, dtype)
    o = capi.default_options(dtype, need_avg_sep=True, isa=H.ref_isa(), periodic=True, boxsize=L)
    r = capi.call_DD(ref, 1, os.cpu_count() or 4, edges, x, y, z, w1=w, weight_type="pair_product", options=o)
    g = T.DD(1, 4, edges, x, y, z, weights1=w, weight_type="pair_product", periodic=True, boxsize=L, output_ravg=True)
    assert np.array_equal(g["npairs"], r["npairs"])
    _close(g["ravg"], r["ravg"], 1e-10 if dtype == np.float64 else 1e-4, "ravg")


# ---- the fast kernel (pairs_fast.cu): 1-D statistics without per-pair sums --------------------------

def _force(kind):
    from corrfunc_b200 import _lib

    _lib.load().cfb_force_kernel(kind)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("periodic", [True, False])
@pytest.mark.parametrize("autocorr", [1, 0])
def test_fast_DD_vs_oracle(dtype, periodic, autocorr, edges):
    """No ravg / weights -> the level-counting kernel; bit-exact npairs vs the oracle, and the generic
    per-pair kernel must agree with it."""
    from corrfunc_b200 import _lib

    T = _theory()
    L, N = 420.0, 40000
    x, y, z, _ = H.box_points(61, N, L, dtype)
    x2, y2, z2, _ = H.box_points(62, N // 2, L, dtype)
    kw = dict(periodic=periodic, boxsize=L)
    okw = dict(periodic=periodic, boxsize=L, autocorr=bool(autocorr))
    if not autocorr:
        kw.update(X2=x2, Y2=y2, Z2=z2)
        okw.update(X2=x2, Y2=y2, Z2=z2)
    got = T.DD(autocorr, 4, edges, x, y, z, **kw)
    assert _lib.last_stats()["kernel_kind"] == 1
    ref = H.oracle_theory("DD", x, y, z, edges, **okw)
    assert np.array_equal(got["npairs"], ref["npairs"])
    _force(0)
    try:
        gen = T.DD(autocorr, 4, edges, x, y, z, **kw)
        assert _lib.last_stats()["kernel_kind"] in (0, 2)  # not the fast kernel: per-pair-sum (2) or legacy generic (0)
    finally:
        _force(-1)
    assert np.array_equal(got["npairs"], gen["npairs"])


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("occ", [0, 12, 40, 300])
def test_fast_dense_cells_and_subdivision(dtype, occ):
    """Dense boxes: many tiles per cell (same-cell rectangle + diagonal jobs), chunked secondaries,
    analytic (no-level) cell pairs on fine lattices; rmin = 0 bins included."""
    from corrfunc_b200 import _lib

    T = _theory()
    L, N = 60.0, 50000
    x, y, z, _ = H.box_points(63, N, L, dtype)
    lib = _lib.load()
    for bins in (np.logspace(-1, np.log10(14.0), 13), np.linspace(0.0, 12.0, 25), np.array([2.0, 9.0])):
        ref = H.oracle_theory("DD", x, y, z, bins, periodic=True, boxsize=L)
        lib.cfb_set_target_occupancy(occ)
        try:
            got = T.DD(1, 4, bins, x, y, z, periodic=True, boxsize=L)
            st = _lib.last_stats()
        finally:
            lib.cfb_set_target_occupancy(0)
        assert st["kernel_kind"] == 1
        assert np.array_equal(got["npairs"], ref["npairs"]), st


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_fast_xi_wp_vs_oracle(dtype, edges):
    from corrfunc_b200 import _lib

    T = _theory()
    L, N, pimax = 420.0, 60000, 40.0
    x, y, z, _ = H.box_points(64, N, L, dtype)
    got = T.xi(L, 4, edges, x, y, z)
    assert _lib.last_stats()["kernel_kind"] == 1
    ref = H.oracle_theory("xi", x, y, z, edges, boxsize=L)
    assert np.array_equal(got["npairs"], ref["npairs"])
    assert np.allclose(got["xi"], ref["cf"], rtol=1e-6 if dtype == np.float64 else 1e-2, atol=1e-9 if dtype == np.float64 else 1e-3)
    for pm in (pimax, 7.5):
        got = T.wp(L, pm, 4, edges, x, y, z)
        assert _lib.last_stats()["kernel_kind"] == 1
        ref = H.oracle_theory("wp", x, y, z, edges, boxsize=L, pimax=pm)
        assert np.array_equal(got["npairs"], ref["npairs"])
    # a dense small box: fine lattice + pi cut through the cells
    L2 = 50.0
    x, y, z, _ = H.box_points(65, N, L2, dtype)
    bins = np.logspace(-1, np.log10(10.0), 10)
    got = T.wp(L2, 6.0, 4, bins, x, y, z)
    ref = H.oracle_theory("wp", x, y, z, bins, boxsize=L2, pimax=6.0)
    assert np.array_equal(got["npairs"], ref["npairs"])


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("autocorr", [1, 0])
@pytest.mark.parametrize("link", [(1, 1), (1, 0), (0, 0)])
def test_fast_DDtheta_vs_oracle(dtype, autocorr, link):
    from corrfunc_b200 import _lib
    from corrfunc_b200.mocks import DDtheta_mocks

    ra1, dec1 = H.sphere_points(15, 40000, dtype)
    ra2, dec2 = H.sphere_points(16, 25000, dtype)
    tb = np.logspace(np.log10(0.05), 1, 16)
    kw = dict(link_in_dec=bool(link[0]), link_in_ra=bool(link[1]))
    okw = dict(link_in_dec=link[0], link_in_ra=link[1], autocorr=bool(autocorr))
    if not autocorr:
        kw.update(RA2=ra2, DEC2=dec2)
        okw.update(RA2=ra2, DEC2=dec2)
    got = DDtheta_mocks(autocorr, 4, tb, ra1, dec1, **kw)
    assert _lib.last_stats()["kernel_kind"] == 1
    ref = H.oracle_theta(ra1, dec1, tb, **okw)
    assert np.array_equal(got["npairs"], ref["npairs"])


def test_fast_DDtheta_reference_golden_counts():
    """npairs of the reference's Mr19 DDtheta golden file through the fast kernel (no weights/thetaavg)."""
    from corrfunc_b200.mocks import DDtheta_mocks

    ra, dec, _ = H.load_mr19_mock()
    bins = H.load_bins_file("angular_bins.txt")
    gold = H.load_wtheta_golden()
    r = DDtheta_mocks(1, 4, bins, ra, dec)
    assert np.array_equal(r["npairs"], gold["npairs"])


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_fast_against_reference_golden_counts(dtype):
    """npairs of the committed reference outputs, through the fast kernel."""
    T = _theory()
    g = np.load(os.path.join(H.GOLDEN, "ref_synthetic_%s.npz" % np.dtype(dtype).name))
    seed, N, L, edges = int(g["seed"]), int(g["N"]), float(g["L"]), g["edges"]
    x, y, z, _ = H.box_points(seed, N, L, dtype)
    x2, y2, z2, _ = H.box_points(seed + 1, N // 2, L, dtype)
    for periodic in (True, False):
        p = "per" if periodic else "nonper"
        r = T.DD(1, 4, edges, x, y, z, periodic=periodic, boxsize=L)
        assert np.array_equal(r["npairs"], g["DD_auto_%s__npairs" % p])
        r = T.DD(0, 4, edges, x, y, z, periodic=periodic, boxsize=L, X2=x2, Y2=y2, Z2=z2)
        assert np.array_equal(r["npairs"], g["DD_cross_%s__npairs" % p])
    r = T.xi(L, 4, edges, x, y, z)
    assert np.array_equal(r["npairs"], g["xi__npairs"])
    r = T.wp(L, 40.0, 4, edges, x, y, z)
    assert np.array_equal(r["npairs"], g["wp__npairs"])


# ------------------------------------------------------------------------------------------------
# The generic kernel keeps its per-block sums as 96-bit fixed point (three native 32-bit shared-memory
# atomics with carries, scaled by the bin's upper edge / the largest weight product): signed weights,
# weights of very different magnitude and a first bin that starts at zero must still meet the tolerance.
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("kind", ["signed", "wide", "large"])
def test_weighted_sums_fixed_point(dtype, kind):
    T = _theory()
    L, N = 100.0, 20000
    x, y, z, w = H.box_points(31, N, L, dtype)
    rng = np.random.default_rng(32)
    if kind == "signed":
        w = (rng.random(N) - 0.5).astype(dtype)
    elif kind == "wide":
        w = (10.0 ** rng.uniform(-6, 3, N)).astype(dtype)
    else:
        w = (1.0e6 * (1.0 + rng.random(N))).astype(dtype)
    edges = np.concatenate([[0.0], np.logspace(-1, np.log10(12.0), 9)])  # rmin = 0: self pairs enter the first bin
    got = T.DD(1, 2, edges, x, y, z, weights1=w, weight_type="pair_product", periodic=True, boxsize=L, output_ravg=True)
    ref = H.oracle_theory("DD", x, y, z, edges, w1=w, weight_type="pair_product", periodic=True, boxsize=L, need_avg=True)
    assert np.array_equal(got["npairs"], ref["npairs"])
    _close(got["ravg"], ref["ravg"], TOL[dtype], "ravg")
    # signed weights cancel in the sum: compare against the scale of the summed magnitudes
    wsum_got = got["weightavg"] * got["npairs"]
    wsum_ref = ref["weightavg"] * ref["npairs"]
    scale = np.abs(w).astype(np.float64).max() ** 2 * np.maximum(ref["npairs"], 1)
    assert np.max(np.abs(wsum_got - wsum_ref) / scale) <= TOL[dtype], kind
    if kind != "signed":
        _close(got["weightavg"], ref["weightavg"], TOL[dtype], "weightavg")
    got = T.DDsmu(1, 2, edges[1:], 1.0, 7, x, y, z, weights1=w, weight_type="pair_product", periodic=True, boxsize=L,
                  output_savg=True)
    ref = H.oracle_theory("DDsmu", x, y, z, edges[1:], w1=w, weight_type="pair_product", periodic=True, boxsize=L,
                          need_avg=True, mu_max=1.0, nmu_bins=7)
    assert np.array_equal(got["npairs"], ref["npairs"].ravel())
    _close(got["savg"], ref["ravg"].ravel(), TOL[dtype], "savg")


def test_precision_suffixed_entry_points():
    """countpairs_float / countpairs_xi_double ... (the reference's *_impl.h.src prototypes) take typed pointers
    and ignore options->float_type."""
    import ctypes as C

    from corrfunc_b200 import _capi, _lib

    lib = _lib.load()
    _capi._declare(lib)
    L, N = 150.0, 20000
    bins = np.logspace(-1, np.log10(10.0), 9)
    for dtype, suf in ((np.float32, "float"), (np.float64, "double")):
        x, y, z, _ = H.box_points(41, N, L, dtype)
        want = _capi.call_DD(lib, 1, 1, bins, x, y, z, options=_capi.default_options(dtype, periodic=True, boxsize=L))
        o = _capi.default_options(dtype, periodic=True, boxsize=L)
        o.float_type = 12 - o.float_type  # deliberately the other width: the typed entry point must not look at it
        e, keep = _capi.make_extra(None, None, None, dtype)
        r = _capi.ResultsDD()
        fn = getattr(lib, "countpairs_" + suf)
        fn.restype = C.c_int
        with _capi.binfile_for(bins) as bf:
            p = [C.c_void_p(a.ctypes.data) for a in (x, y, z)]
            st = fn(C.c_int64(N), p[0], p[1], p[2], C.c_int64(N), p[0], p[1], p[2], C.c_int(1), C.c_int(1), bf,
                    C.byref(r), C.byref(o), C.byref(e))
        assert st == 0
        got = np.ctypeslib.as_array(r.npairs, shape=(r.nbin,)).astype(np.uint64)[1:].copy()
        lib.free_results(C.byref(r))
        assert np.array_equal(got, want["npairs"])
        assert o.float_type == 12 - np.dtype(dtype).itemsize  # restored


@pytest.mark.parametrize("stat", ["xi", "DDsmu"])
@pytest.mark.parametrize("nranks", [2, 3])
def test_rank_sharding_sums_to_the_full_count(stat, nranks):
    """Multi-GPU sharding emulated in one process: the library is run once per rank (each run sorts the particles
    again, in another order inside the cells, exactly like separate processes do) and the partial histograms are
    summed through the reduce hook.  Cells hold ~140 particles here, i.e. two primary tiles each: npairs must not
    depend on the number of ranks."""
    from corrfunc_b200 import _lib

    T = _theory()
    L, N = 100.0, 560000
    x, y, z, w = H.box_points(51, N, L, np.float32)
    bins = np.logspace(-1, 1, 11)

    def run():
        if stat == "xi":
            return T.xi(L, 1, bins, x, y, z)["npairs"]
        return T.DDsmu(1, 1, bins, 1.0, 5, x, y, z, periodic=True, boxsize=L, weights1=w, weight_type="pair_product",
                       output_savg=True)["npairs"]

    full = run()
    acc = {}

    def hook(n, s, ww):
        if acc:
            n += acc["n"]
            s += acc["s"]
            ww += acc["w"]
        acc["n"], acc["s"], acc["w"] = n.copy(), s.copy(), ww.copy()

    try:
        for r in range(nranks):
            _lib.set_shard(r, nranks, hook)
            got = run()  # complete after the last rank's hook
    finally:
        _lib.set_shard(0, 1, None)
    assert np.array_equal(got, full)
    if stat == "xi":
        ref = H.oracle_theory("xi", x, y, z, bins, boxsize=L)
        assert np.array_equal(full, ref["npairs"])


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("autocorr", [1, 0])
@pytest.mark.parametrize("link", [(1, 1), (1, 0), (0, 0)])
@pytest.mark.parametrize("occ", [6, 25])
def test_DDtheta_refined_lattice(dtype, autocorr, link, occ):
    """The device-only sub x sub refinement of the reference RA/DEC cells (cfb_theta_subdivision) must not change a
    single count: a small target occupancy forces sub > 1 at test sizes, for both kernels (the plain counts take the
    fast kernel, weights + thetaavg the generic one), every linking mode, auto and cross."""
    from corrfunc_b200 import _lib
    from corrfunc_b200.mocks import DDtheta_mocks

    ra1, dec1 = H.sphere_points(15, 40000, dtype)
    ra2, dec2 = H.sphere_points(16, 25000, dtype)
    w1 = (1.0 - np.random.default_rng(17).random(ra1.size)).astype(dtype)
    w2 = (1.0 - np.random.default_rng(18).random(ra2.size)).astype(dtype)
    tb = np.logspace(np.log10(0.05), 1, 16)
    lk = dict(link_in_dec=bool(link[0]), link_in_ra=bool(link[1]))
    kw_plain, kw_full = dict(lk), dict(lk, weights1=w1, weight_type="pair_product", output_thetaavg=True)
    okw = dict(w1=w1, weight_type="pair_product", need_avg=True, link_in_dec=link[0], link_in_ra=link[1], autocorr=bool(autocorr))
    if not autocorr:
        kw_plain.update(RA2=ra2, DEC2=dec2)
        kw_full.update(RA2=ra2, DEC2=dec2, weights2=w2)
        okw.update(RA2=ra2, DEC2=dec2, w2=w2)
    ref = H.oracle_theta(ra1, dec1, tb, **okw)
    lib = _lib.load()
    lib.cfb_set_target_occupancy(occ)
    try:
        got = DDtheta_mocks(autocorr, 2, tb, ra1, dec1, **kw_plain)
        st = _lib.last_stats()
        assert st["kernel_kind"] == 1 and st["n_cells"] >= 1
        assert np.array_equal(got["npairs"], ref["npairs"])
        got = DDtheta_mocks(autocorr, 2, tb, ra1, dec1, **kw_full)
        assert _lib.last_stats()["kernel_kind"] in (0, 2)  # not the fast kernel: per-pair-sum (2) or legacy generic (0)
    finally:
        lib.cfb_set_target_occupancy(0)
    assert np.array_equal(got["npairs"], ref["npairs"])
    _close(got["thetaavg"], ref["ravg"], 1e-9 if dtype == np.float64 else 1e-4, "thetaavg")
    _close(got["weightavg"], ref["weightavg"], TOL[dtype], "weightavg")


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_DDrppi_histogram_too_large_for_shared_memory(dtype):
    """(nbin+1)*(npibin+1) = 27*251 slots x (count + two 96-bit sums) exceed the block's shared-memory budget: the
    generic kernel then accumulates straight into global memory (native 64-bit / double atomics)."""
    T = _theory()
    L, N, pimax = 600.0, 30000, 250.0
    x, y, z, w = H.box_points(71, N, L, dtype)
    bins = np.logspace(-0.5, np.log10(40.0), 27)
    got = T.DDrppi(1, 2, pimax, bins, x, y, z, weights1=w, weight_type="pair_product", periodic=True, boxsize=L,
                   output_rpavg=True)
    ref = H.oracle_theory("DDrppi", x, y, z, bins, w1=w, weight_type="pair_product", periodic=True, boxsize=L,
                          need_avg=True, pimax=pimax)
    assert np.array_equal(got["npairs"], ref["npairs"].ravel())
    _close(got["rpavg"], ref["ravg"].ravel(), TOL[dtype], "rpavg")
    _close(got["weightavg"], ref["weightavg"].ravel(), TOL[dtype], "weightavg")


# ---- BASELINE configs at FULL size against the unmodified reference ----------------------------------

def _run_config(lib, name):
    """One BASELINE config on bench.py's own synthetic input, through the C ABI (same helper that
    tests/golden/make_golden_fullsize.py used to drive oracle/_ref)."""
    import bench
    from corrfunc_b200 import _capi

    cfg = bench.config_by_name(name)
    dtype = np.float32 if cfg["dtype"] == "f32" else np.float64
    bins = bench.make_bins(cfg["bins"])
    pts = bench.gen_points(cfg, cfg["N"], dtype)
    o = _capi.default_options(dtype, periodic=True, need_avg_sep=bool(cfg.get("avg")),
                              boxsize=cfg["L"] if cfg["L"] > 0 else None)
    wt = "pair_product" if cfg.get("weights") else None
    st = cfg["stat"]
    if st == "xi":
        return _capi.call_xi(lib, cfg["L"], 1, bins, pts["x"], pts["y"], pts["z"], options=o)
    if st == "DD":
        return _capi.call_DD(lib, 1, 1, bins, pts["x"], pts["y"], pts["z"], options=o)
    if st == "wp":
        return _capi.call_wp(lib, cfg["L"], 1, cfg["pimax"], bins, pts["x"], pts["y"], pts["z"], options=o)
    if st == "DDrppi":
        return _capi.call_DDrppi(lib, 1, 1, cfg["pimax"], bins, pts["x"], pts["y"], pts["z"], options=o)
    if st == "DDsmu":
        return _capi.call_DDsmu(lib, 1, 1, bins, cfg["mu_max"], cfg["nmu"], pts["x"], pts["y"], pts["z"],
                                w1=pts.get("w"), weight_type=wt, options=o)
    return _capi.call_DDtheta(lib, 0, 1, bins, pts["ra"], pts["dec"], RA2=pts["ra2"], DEC2=pts["dec2"], options=o)


@pytest.mark.parametrize("stat", ["DD", "xi"])
def test_float_z_window_case(stat):
    """The two float32 inputs on which the reference's z-window pruning loses one pair of the last bin
    (tests/golden/make_golden_float_window.py; mechanism pinned by the oracle's literal mode in
    tests/test_cpu_oracle.py): the GPU evaluates every candidate pair, so GPU = reference + 2 in the last bin and
    is identical everywhere else."""
    T = _theory()
    ref = np.load(os.path.join(H.GOLDEN, "ref_float_window.npz"))[stat].astype(np.int64)
    x, y, z, L, edges = H.float_window_case(stat)
    r = T.xi(L, 1, edges, x, y, z) if stat == "xi" else T.DD(1, 1, edges, x, y, z, periodic=True, boxsize=L)
    d = r["npairs"].astype(np.int64) - ref
    assert not d[:-1].any() and d[-1] == 2, d


# c5dsd10M: config 5 in DOUBLE at the same density (10 M points) -- the double kernel is another code path
# (compare-and-count instead of the packed sign-bit counters); c5DDsd10M: the same points through DD(autocorr=1, periodic)
# c5d: config 5 itself in double, 100 M points (the reference needed 6151 s on 5 cores for this golden)
FULL_SIZE_VERIFIED = ["c1", "c2", "c2wp32", "c2rppi", "c2rppi32", "c3", "c4", "c5sd10M", "c5dsd10M", "c5DDsd10M", "c5", "c5d"]


@pytest.mark.parametrize("name", FULL_SIZE_VERIFIED)
def test_full_size_config_vs_reference_golden(name):
    _check_full_size(name)


def _check_full_size(name):
    """BASELINE configs 1-4 at their full sizes (1.2M / 10M / 2M+2M points): npairs bit-exact against the
    committed outputs of the UNMODIFIED reference (oracle/_ref, AVX-512F kernels) on the same seeded inputs
    (tests/golden/make_golden_fullsize.py); ravg / weightavg of config 3 within 1e-10 relative.

    c1f32 / c3f32 / c4f32 are configs 1, 3 and 4 on float32 inputs (the CPU oracle agrees with the reference on all
    three at full size: no early-exit effect there).
    Config 5 itself (xi, 100 M points, float; the reference needs 25 minutes on 8 cores for it): 29 of the 30 bins are
    bit-exact and the last one holds 4 more (2 unordered pairs of 8.8e12) than the reference, whose float z-window
    never visits them (test_float_z_window_case).

    Float: wp is bit-exact too.  DDrppi in float differs by exactly the pairs the reference's AVX-512 kernel
    LOSES to a rounding quirk of its early exit (148 unordered pairs of 1.5e9; mechanism and count pinned by the
    oracle's literal mode, tests/golden/make_golden_rppi32_dropped.py, tests/test_cpu_oracle.py): the GPU counts
    them, so GPU == reference + dropped, bin by bin."""
    from corrfunc_b200 import _lib

    g = np.load(os.path.join(H.GOLDEN, "ref_fullsize_%s.npz" % name))
    want = g["npairs"].astype(np.int64)
    if name == "c5":
        want[-1] += 4  # the stated float-path difference of config 5: last bin only
    if name == "c2rppi32":
        dropped = np.load(os.path.join(H.GOLDEN, "ref_fullsize_c2rppi32_dropped.npz"))["dropped"].astype(np.int64)
        assert dropped.sum() == 296 and np.count_nonzero(dropped) == 82  # the stated float-path difference
        want = want + dropped
    r = _run_config(_lib.load(), name)
    got = np.asarray(r["npairs"], dtype=np.int64).reshape(want.shape)
    nflip = int(np.count_nonzero(got != want))
    assert nflip == 0, "%s: %d bins differ from the reference, max |diff| %d" % (name, nflip, int(np.abs(got - want).max()))
    for k in ("ravg", "weightavg"):
        # float goldens: npairs only -- the reference sums separations in float, which at 1e8 pairs per bin has
        # lost most of its digits, so its float averages are not a standard to hold the GPU to
        if k in g.files and not name.endswith("f32"):
            _close(np.asarray(r[k]).reshape(g[k].shape), g[k], 1e-10, k)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("autocorr", [1, 0])
@pytest.mark.parametrize("occ", [0, 8])
def test_DDtheta_fast_acos_vs_oracle(dtype, autocorr, occ):
    """options.fast_acos = 1: thetaavg from the degree-8 polynomial of utils/fast_acos.h:57-101 (used by the reference at
    countpairs_theta_mocks_kernels.c.src:153,490,822) instead of libm's acos.  Same counts, and averages equal to the
    oracle's run with the same flag -- and measurably different from the libm run, so the flag is known to reach the
    kernel.  occ = 8 forces the sub x sub refinement of the RA/DEC lattice."""
    from corrfunc_b200 import _lib
    from corrfunc_b200.mocks import DDtheta_mocks

    ra1, dec1 = H.sphere_points(25, 30000, dtype)
    ra2, dec2 = H.sphere_points(26, 20000, dtype)
    w1 = (1.0 - np.random.default_rng(27).random(ra1.size)).astype(dtype)
    w2 = (1.0 - np.random.default_rng(28).random(ra2.size)).astype(dtype)
    tb = np.logspace(np.log10(0.05), 1, 16)
    kw = dict(weights1=w1, weight_type="pair_product", output_thetaavg=True)
    okw = dict(w1=w1, weight_type="pair_product", need_avg=True, autocorr=bool(autocorr))
    if not autocorr:
        kw.update(RA2=ra2, DEC2=dec2, weights2=w2)
        okw.update(RA2=ra2, DEC2=dec2, w2=w2)
    lib = _lib.load()
    lib.cfb_set_target_occupancy(occ)
    try:
        fast = DDtheta_mocks(autocorr, 2, tb, ra1, dec1, fast_acos=True, **kw)
        slow = DDtheta_mocks(autocorr, 2, tb, ra1, dec1, fast_acos=False, **kw)
    finally:
        lib.cfb_set_target_occupancy(0)
    ref = H.oracle_theta(ra1, dec1, tb, fast_acos=True, **okw)
    assert np.array_equal(fast["npairs"], ref["npairs"]) and np.array_equal(slow["npairs"], ref["npairs"])
    _close(fast["thetaavg"], ref["ravg"], 1e-9 if dtype == np.float64 else 1e-4, "thetaavg (fast_acos)")
    _close(fast["weightavg"], ref["weightavg"], TOL[dtype], "weightavg")
    if dtype == np.float64:
        # the polynomial is good to 3.7e-9 rad, libm to 1e-16: the two runs differ, but by less than 1e-6 degrees
        d = np.abs(fast["thetaavg"] - slow["thetaavg"])
        assert d.max() > 0 and d.max() < 1e-6


@pytest.mark.parametrize("stat", ["DD", "xi"])
def test_device_pointers_including_weights(stat):
    """Positions AND weights handed over as device memory (cfb_upload borrows device pointers; INTEGRATION.md tells
    callers with resident catalogues to do that).  The host epilogue needs the weights for the self-pair term (bins
    starting at 0) and for the weight sums of xi: it reads them back (cfb_is_device_ptr / cfb_copy_to_host) instead of
    dereferencing a device pointer.  Same results as from host arrays."""
    import ctypes as C

    import torch
    from corrfunc_b200 import _capi, _lib

    lib = _lib.load()
    _capi._declare(lib)
    L, N = 150.0, 60000
    x, y, z, w = H.box_points(77, N, L, np.float64)
    bins = np.linspace(0.0, 12.0, 9)  # rmin = 0: the self pairs and their w*w enter the first bin
    o = _capi.default_options(np.float64, periodic=True, need_avg_sep=True, boxsize=L)
    if stat == "DD":
        host = _capi.call_DD(lib, 1, 1, bins, x, y, z, w1=w, weight_type="pair_product", options=o)
    else:
        host = _capi.call_xi(lib, L, 1, bins, x, y, z, w=w, weight_type="pair_product", options=o)
    d = [torch.from_numpy(a).cuda() for a in (x, y, z, w)]
    P = [C.c_void_p(t.data_ptr()) for t in d]
    e = _capi.ExtraOptions()
    e.weight_method = _capi.WEIGHT_PAIR_PRODUCT
    for ws in (e.weights0, e.weights1):
        ws.num_weights = 1
        ws.weights[0] = d[3].data_ptr()
    o = _capi.default_options(np.float64, periodic=True, need_avg_sep=True, boxsize=L)
    with _capi.binfile_for(bins) as bf:
        if stat == "DD":
            r = _capi.ResultsDD()
            st = lib.countpairs(N, P[0], P[1], P[2], N, P[0], P[1], P[2], 1, 1, bf, C.byref(r), C.byref(o), C.byref(e))
        else:
            r = _capi.ResultsXi()
            st = lib.countpairs_xi(N, P[0], P[1], P[2], L, 1, bf, C.byref(r), C.byref(o), C.byref(e))
    assert st == 0
    n = r.nbin
    got_np = np.ctypeslib.as_array(r.npairs, shape=(n,)).copy()[1:]
    got_w = np.ctypeslib.as_array(r.weightavg, shape=(n,)).copy()[1:]
    assert np.array_equal(got_np, host["npairs"])
    _close(got_w, host["weightavg"], 1e-12, "weightavg from device weights")
    if stat == "xi":
        # xi = DD/RR - 1 is ~1e-5 here: the 1e-16 run-to-run spread of the weight sums shows up as 1e-11 relative
        assert np.allclose(np.ctypeslib.as_array(r.xi, shape=(n,)).copy()[1:], host["cf"], rtol=1e-9, atol=1e-13)
        lib.free_results_xi(C.byref(r))
    else:
        lib.free_results(C.byref(r))


@pytest.mark.parametrize("stat", ["DD", "DDrppi", "DDsmu"])
def test_DD_DR_RR_in_one_context(stat):
    """corrfunc_b200.workflow.DD_DR_RR: data and randoms uploaded once (catalogue cache), three counts -- identical to
    three separate calls, with the uploads of the second and third count served from the cache; the Landy-Szalay
    estimate on top equals the reference's converter on the separate counts (Corrfunc/utils.py:27-165)."""
    from corrfunc_b200 import _lib, workflow
    from corrfunc_b200.utils import convert_3d_counts_to_cf

    T = _theory()
    L = 200.0
    x, y, z, w = H.box_points(81, 50000, L, np.float64)
    rx, ry, rz, rw = H.box_points(82, 120000, L, np.float64)
    bins = np.logspace(-0.3, np.log10(18.0), 12)
    kw = dict(periodic=True, boxsize=L, weight_type="pair_product")
    extra = dict(DD={}, DDrppi=dict(pimax=20.0), DDsmu=dict(mu_max=1.0, nmu_bins=8))[stat]
    lib = _lib.load()
    hits0 = lib.cfb_catalog_cache_hits()
    dd, dr, rr = workflow.DD_DR_RR(2, bins, x, y, z, rx, ry, rz, weights=w, rweights=rw, stat=stat, **extra, **kw)
    # DD(D) uploads D; DR finds D, uploads R; RR finds R (in the other slot): 2 hits
    assert lib.cfb_catalog_cache_hits() - hits0 == 2
    if stat == "DD":
        fn = lambda a, A, wa, **k: T.DD(a, 2, bins, A[0], A[1], A[2], weights1=wa, **k, **kw)
    elif stat == "DDrppi":
        fn = lambda a, A, wa, **k: T.DDrppi(a, 2, 20.0, bins, A[0], A[1], A[2], weights1=wa, **k, **kw)
    else:
        fn = lambda a, A, wa, **k: T.DDsmu(a, 2, bins, 1.0, 8, A[0], A[1], A[2], weights1=wa, **k, **kw)
    dd2 = fn(1, (x, y, z), w)
    dr2 = fn(0, (x, y, z), w, X2=rx, Y2=ry, Z2=rz, weights2=rw)
    rr2 = fn(1, (rx, ry, rz), rw)
    for a, b in ((dd, dd2), (dr, dr2), (rr, rr2)):
        assert np.array_equal(a["npairs"], b["npairs"])
        _close(a["weightavg"], b["weightavg"], 1e-12, "weightavg")
    if stat == "DD":
        cf = workflow.xi_from_catalogs(2, bins, x, y, z, rx, ry, rz, periodic=True, boxsize=L)
        cf2 = convert_3d_counts_to_cf(x.size, x.size, rx.size, rx.size, T.DD(1, 2, bins, x, y, z, periodic=True, boxsize=L),
                                      T.DD(0, 2, bins, x, y, z, X2=rx, Y2=ry, Z2=rz, periodic=True, boxsize=L),
                                      T.DD(0, 2, bins, x, y, z, X2=rx, Y2=ry, Z2=rz, periodic=True, boxsize=L),
                                      T.DD(1, 2, bins, rx, ry, rz, periodic=True, boxsize=L))
        assert np.allclose(cf, cf2, rtol=0, atol=0)
    # outside the context every call uploads again
    h1 = lib.cfb_catalog_cache_hits()
    fn(1, (x, y, z), w)
    fn(1, (x, y, z), w)
    assert lib.cfb_catalog_cache_hits() == h1


def test_two_pass_scatter_forced_on_small_sets():
    """Sets of 4 M points and more are sorted by the two-pass scatter (gridlink.cu: k_partition / k_place); the
    full-size goldens cover it for box lattices.  Here the threshold is lowered to zero in a fresh process, so the
    driver's smoke cases -- weighted DD + ravg in double, xi in float, DDtheta on the RA/DEC lattice, each against the
    oracle -- run through it on small inputs as well (records with weights, both precisions, the theta lattice)."""
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, CORRFUNC_B200_SORT2_MIN="0")
    r = subprocess.run([sys.executable, "-c", "import __graft_entry__ as g; g.smoke()"], cwd=root, env=env,
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "smoke OK" in r.stdout
