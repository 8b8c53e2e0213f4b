"""CPU tests of the checker itself: the oracle restatement vs (a) the reference's golden file,
(b) committed outputs of the unmodified reference on seeded inputs, (c) the reference's data-free
known-answer tests, (d) the live unmodified reference when oracle/_ref is present."""
import os

import numpy as np
import pytest

import harness as H


def test_oracle_matches_reference_golden_wtheta():
    """mocks/tests/Mr19_mock_wtheta.DD: all 20 npairs exact, averages to the file's print precision."""
    ra, dec, w = H.load_mr19_mock()
    bins = H.load_bins_file("angular_bins.txt")
    gold = H.load_wtheta_golden()
    a = H.oracle_theta(ra, dec, bins, w1=w, weight_type="pair_product", need_avg=True)
    assert np.array_equal(a["npairs"], gold["npairs"])
    assert np.allclose(a["ravg"], gold["ravg"], atol=1e-8, rtol=1e-6)
    assert np.allclose(a["weightavg"], gold["weightavg"], atol=1e-8, rtol=1e-6)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_oracle_matches_committed_reference_outputs(dtype):
    g = np.load(os.path.join(H.GOLDEN, "ref_synthetic_%s.npz" % np.dtype(dtype).name))
    seed, N, L, edges = int(g["seed"]), int(g["N"]), float(g["L"]), g["edges"]
    x, y, z, w = H.box_points(seed, N, L, dtype)
    x2, y2, z2, w2 = H.box_points(seed + 1, N // 2, L, dtype)
    tol = 1e-10 if dtype == np.float64 else 1e-4
    kw = dict(w1=w, weight_type="pair_product", need_avg=True, boxsize=L)
    for periodic in (True, False):
        p = "per" if periodic else "nonper"
        a = H.oracle_theory("DD", x, y, z, edges, periodic=periodic, **kw)
        assert np.array_equal(a["npairs"], g["DD_auto_%s__npairs" % p])
        assert np.allclose(a["ravg"], g["DD_auto_%s__ravg" % p], rtol=tol)
        assert np.allclose(a["weightavg"], g["DD_auto_%s__weightavg" % p], rtol=tol)
        a = H.oracle_theory("DD", x, y, z, edges, periodic=periodic, autocorr=False, X2=x2, Y2=y2, Z2=z2, w2=w2, **kw)
        assert np.array_equal(a["npairs"], g["DD_cross_%s__npairs" % p])
        a = H.oracle_theory("DDrppi", x, y, z, edges, periodic=periodic, pimax=40.0, **kw)
        assert np.array_equal(a["npairs"], g["DDrppi_auto_%s__npairs" % p])
        assert np.allclose(a["ravg"], g["DDrppi_auto_%s__ravg" % p], rtol=tol)
        a = H.oracle_theory("DDsmu", x, y, z, edges, periodic=periodic, mu_max=0.5, nmu_bins=10, **kw)
        assert np.array_equal(a["npairs"], g["DDsmu_auto_%s__npairs" % p])
        assert np.allclose(a["ravg"], g["DDsmu_auto_%s__ravg" % p], rtol=tol)
    a = H.oracle_theory("xi", x, y, z, edges, **kw)
    assert np.array_equal(a["npairs"], g["xi__npairs"])
    assert np.allclose(a["cf"], g["xi__cf"], rtol=1e-6 if dtype == np.float64 else 2e-2, atol=1e-9 if dtype == np.float64 else 1e-3)
    a = H.oracle_theory("wp", x, y, z, edges, pimax=40.0, **kw)
    assert np.array_equal(a["npairs"], g["wp__npairs"])


@pytest.mark.parametrize("N", [1, 2])
def test_oracle_narrow_extent(N):  # Corrfunc/tests/test_theory.py:115-147
    pos = np.array([[0.0, 0.0], [0.0, 0.0], [0.0, 0.5]]) if N == 2 else np.array([[0.1], [0.2], [0.3]])
    a = H.oracle_theory("DD", pos[0], pos[1], pos[2], [0.2, 0.6, 1.0], periodic=True, boxsize=(3.0, 3.0, 3.0))
    assert np.all(a["npairs"] == ([2, 0] if N == 2 else [0, 0]))


@pytest.mark.parametrize("autocorr", [0, 1])
@pytest.mark.parametrize("binref", [1, 2, 3])
@pytest.mark.parametrize("maxcells", [1, 2, 3])
def test_oracle_duplicate_cellpairs(autocorr, binref, maxcells):  # test_theory.py:150-194
    boxsize = 432.0
    kw = dict(periodic=True, boxsize=boxsize, refine=(binref,) * 3, custom_refine=binref != 2 or True, max_cells=maxcells,
              autocorr=bool(autocorr))
    pos = np.array([[0.02, 0.98], [0.0, 0.0], [0.0, 0.0]]) * boxsize
    x2 = {} if autocorr else dict(X2=pos[0], Y2=pos[1], Z2=pos[2])
    a = H.oracle_theory("DD", pos[0], pos[1], pos[2], np.array([0.01, 0.4]) * boxsize, **kw, **x2)
    assert np.all(a["npairs"] == [2])
    pos = np.array([[0.0, 0.0], [0.0, 0.0], [0.0, 0.48]]) * boxsize
    x2 = {} if autocorr else dict(X2=pos[0], Y2=pos[1], Z2=pos[2])
    a = H.oracle_theory("DD", pos[0], pos[1], pos[2], np.array([0.2, 0.3, 0.49]) * boxsize, **kw, **x2)
    assert np.all(a["npairs"] == [0, 2])


@pytest.mark.parametrize("autocorr", [0, 1], ids=["cross", "auto"])
@pytest.mark.parametrize("binref", [1, 3], ids=["ref1", "ref3"])
@pytest.mark.parametrize("maxcells", [1, 3], ids=["max1", "max3"])
@pytest.mark.parametrize("boxsize", [123.0, (51.0, 75.0, 123.0)], ids=["iso", "aniso"])
@pytest.mark.parametrize("funcname", ["DD", "DDrppi", "DDsmu"])
@pytest.mark.parametrize("periodic", [False, True], ids=["nowrap", "wrap"])
def test_oracle_brute(autocorr, binref, maxcells, boxsize, funcname, periodic):  # test_theory.py:197-286
    np.random.seed(1234)
    npts, eps = 100, 0.2
    boxsize = np.array(boxsize)
    bins = np.linspace(0.01, 0.49 * boxsize.min(), 20) if periodic else np.linspace(0.01, 2 * boxsize.max(), 20)
    pimax = np.floor(0.49 * boxsize.min())
    mu_max, nmu_bins = 0.5, 10
    pos = np.random.uniform(low=-eps, high=eps, size=(npts, 3)) * boxsize
    pos[npts // 2:] += boxsize / 2.0
    pos %= boxsize
    pdiff = np.abs(pos[:, np.newaxis] - pos)
    if periodic:
        pdiff -= (pdiff >= boxsize / 2) * boxsize
    kw = dict(periodic=periodic, boxsize=boxsize, refine=(binref,) * 3, custom_refine=True, max_cells=maxcells,
              autocorr=bool(autocorr))
    if not autocorr:
        kw.update(X2=pos[:, 0], Y2=pos[:, 1], Z2=pos[:, 2])
    if funcname == "DDrppi":
        kw["pimax"] = pimax
        brute, _, _ = np.histogram2d((pdiff[:, :, :2] ** 2).sum(axis=-1).reshape(-1), np.abs(pdiff[:, :, 2]).reshape(-1),
                                     bins=(bins ** 2, np.linspace(0.0, pimax, int(pimax) + 1)))
    elif funcname == "DDsmu":
        kw.update(mu_max=mu_max, nmu_bins=nmu_bins)
        sdiff = np.sqrt((pdiff ** 2).sum(axis=-1).reshape(-1))
        sdiff[sdiff == 0.0] = np.inf
        brute, _, _ = np.histogram2d(sdiff, np.abs(pdiff[:, :, 2]).reshape(-1) / sdiff,
                                     bins=(bins, np.linspace(0, mu_max, nmu_bins + 1)))
    else:
        brute, _ = np.histogram((pdiff ** 2).sum(axis=-1).reshape(-1), bins=bins ** 2)
    a = H.oracle_theory(funcname, pos[:, 0].copy(), pos[:, 1].copy(), pos[:, 2].copy(), bins, **kw)
    assert np.all(a["npairs"].reshape(brute.shape) == brute)


@pytest.mark.skipif(H.load_ref() is None, reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_oracle_vs_live_reference(dtype):
    from corrfunc_b200 import _capi as capi

    ref, isa = H.load_ref(), H.ref_isa()
    L, N = 300.0, 30000
    x, y, z, w = H.box_points(99, N, L, dtype)
    edges = H.load_bins_file("theory_bins.txt")
    for periodic in (True, False):
        o = capi.default_options(dtype, periodic=periodic, need_avg_sep=True, boxsize=L, isa=isa)
        r = capi.call_DD(ref, 1, 4, edges, x, y, z, w1=w, weight_type="pair_product", options=o)
        a = H.oracle_theory("DD", x, y, z, edges, periodic=periodic, boxsize=L, w1=w, weight_type="pair_product", need_avg=True)
        assert np.array_equal(a["npairs"], r["npairs"])
        o = capi.default_options(dtype, periodic=periodic, need_avg_sep=True, boxsize=L, isa=isa)
        r = capi.call_DDsmu(ref, 1, 4, edges, 1.0, 20, x, y, z, w1=w, weight_type="pair_product", options=o)
        a = H.oracle_theory("DDsmu", x, y, z, edges, periodic=periodic, boxsize=L, w1=w, weight_type="pair_product",
                            need_avg=True, mu_max=1.0, nmu_bins=20)
        assert np.array_equal(a["npairs"], r["npairs"])
    ra, dec = H.sphere_points(3, 20000, dtype)
    tb = np.logspace(-1.3, 1, 12)
    o = capi.default_options(dtype, need_avg_sep=True, isa=isa)
    r = capi.call_DDtheta(ref, 1, 4, tb, ra, dec, options=o)
    a = H.oracle_theta(ra, dec, tb, need_avg=True)
    assert np.array_equal(a["npairs"], r["npairs"])


# ---- the oracle at BASELINE config-2 size against the unmodified reference's float output ------------

def test_oracle_full_size_wp_float_matches_reference():
    """c2wp32 (1.2 M points, float): the oracle's per-pair conditions (fast-forward survivors, signed dz < pimax,
    wp_kernels.c.src:139-142 / 207-221) reproduce the reference's AVX-512 output bit for bit."""
    ref = np.load(os.path.join(H.GOLDEN, "ref_fullsize_c2wp32.npz"))["npairs"]
    assert np.array_equal(H.oracle_config("c2wp32")["npairs"], ref)


def test_oracle_literal_mode_pins_the_float_DDrppi_quirk():
    """c2rppi32: LITERAL mode (z-sorted cells, 16-lane chunks, the |dz|-before-exit order of
    countpairs_rp_pi_kernels.c.src:196-207) is bit-identical to the reference; the default mode differs from it
    by exactly the committed `dropped` pairs (296 ordered pairs in 82 bins, none at |dz| >= 35)."""
    ref = np.load(os.path.join(H.GOLDEN, "ref_fullsize_c2rppi32.npz"))["npairs"].astype(np.int64)
    dropped = np.load(os.path.join(H.GOLDEN, "ref_fullsize_c2rppi32_dropped.npz"))["dropped"].astype(np.int64)
    lit = H.oracle_config("c2rppi32", literal=1)["npairs"].astype(np.int64).reshape(ref.shape)
    assert np.array_equal(lit, ref)
    dflt = H.oracle_config("c2rppi32")["npairs"].astype(np.int64).reshape(ref.shape)
    assert np.array_equal(dflt - ref, dropped)
    assert dropped.sum() == 296 and dropped[:, 35:].sum() == 0


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("stat", ["wp", "DDrppi"])
def test_oracle_literal_equals_default_on_small_inputs(stat, dtype):
    """On inputs where no dz rounds onto -pimax the early exits are pure pruning: both modes agree, weights too."""
    L, N = 420.0, 60000
    x, y, z, w = H.box_points(5, N, L, dtype)
    edges = np.logspace(np.log10(0.1), np.log10(25.0), 15)
    kw = dict(pimax=40.0, periodic=True, boxsize=L, w1=w, weight_type="pair_product", need_avg=True)
    a = H.oracle_theory(stat, x, y, z, edges, **kw)
    lib = H.load_oracle()
    lib.oracle_set_literal_kernels(1)
    try:
        b = H.oracle_theory(stat, x, y, z, edges, **kw)
    finally:
        lib.oracle_set_literal_kernels(0)
    assert np.array_equal(a["npairs"], b["npairs"])
    assert np.allclose(a["ravg"], b["ravg"], rtol=1e-12) and np.allclose(a["weightavg"], b["weightavg"], rtol=1e-12)


# ---- survey geometry: mocks/DDrppi_mocks and mocks/DDsmu_mocks -----------------------

@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_oracle_mocks_match_committed_reference_outputs(dtype):
    """tests/golden/ref_mocks_*.npz: the unmodified reference (AVX-512F kernels) on seeded survey wedges."""
    g = np.load(os.path.join(H.GOLDEN, "ref_mocks_%s.npz" % np.dtype(dtype).name))
    ra, dec, d, w = H.mock_points(int(g["seed"]), int(g["N1"]), dtype)
    ra2, dec2, d2, w2 = H.mock_points(int(g["seed"]) + 1, int(g["N2"]), dtype)
    tol = 1e-10 if dtype == np.float64 else 1e-4  # the reference's float path sums in float
    for autocorr in (1, 0):
        tag = "auto" if autocorr else "cross"
        kw = dict(w1=w, weight_type="pair_product", need_avg=True, autocorr=bool(autocorr), periodic=False)
        if not autocorr:
            kw.update(X2=ra2, Y2=dec2, Z2=d2, w2=w2)
        a = H.oracle_theory("DDrppi_mocks", ra, dec, d, g["edges"], pimax=float(g["pimax"]), **kw)
        assert np.array_equal(a["npairs"], g["DDrppi_mocks_%s__npairs" % tag])
        assert np.allclose(a["ravg"], g["DDrppi_mocks_%s__ravg" % tag], rtol=tol)
        assert np.allclose(a["weightavg"], g["DDrppi_mocks_%s__weightavg" % tag], rtol=tol)
        a = H.oracle_theory("DDsmu_mocks", ra, dec, d, g["edges"], mu_max=float(g["mu_max"]), nmu_bins=int(g["nmu"]), **kw)
        assert np.array_equal(a["npairs"], g["DDsmu_mocks_%s__npairs" % tag])
        assert np.allclose(a["ravg"], g["DDsmu_mocks_%s__ravg" % tag], rtol=tol)
        assert np.allclose(a["weightavg"], g["DDsmu_mocks_%s__weightavg" % tag], rtol=tol)


def test_oracle_mocks_brute_force():
    """Independent numpy restatement of the definitions (pi = |s.l|/|l|, rp^2 = s^2 - pi^2, mu = pi/s with
    l the pair midpoint), all pairs, no lattice: counts agree except where a pair sits within rounding of an edge."""
    ra, dec, d, _ = H.mock_points(21, 1500, np.float64)
    x = d * np.cos(np.radians(dec)) * np.cos(np.radians(ra))
    y = d * np.cos(np.radians(dec)) * np.sin(np.radians(ra))
    z = d * np.sin(np.radians(dec))
    p = np.stack([x, y, z], 1)
    i, j = np.triu_indices(len(p), 1)
    s = p[j] - p[i]
    l = p[j] + p[i]
    s2 = (s * s).sum(1)
    pi2 = (s * l).sum(1) ** 2 / (l * l).sum(1)
    rp2 = s2 - pi2
    edges = np.logspace(np.log10(0.5), np.log10(60.0), 9)
    pimax, mu_max, nmu = 30.0, 0.8, 5
    ok = (pi2 < pimax ** 2) & (rp2 >= edges[0] ** 2) & (rp2 < edges[-1] ** 2)
    want = np.histogram2d(np.sqrt(rp2[ok]), np.sqrt(pi2[ok]), bins=[edges, np.arange(int(pimax) + 1)])[0] * 2
    a = H.oracle_theory("DDrppi_mocks", ra, dec, d, edges, pimax=pimax, periodic=False)
    assert np.abs(a["npairs"].astype(np.int64) - want.astype(np.int64)).sum() <= 4
    mu = np.sqrt(pi2 / s2)
    ok = (mu < mu_max) & (s2 >= edges[0] ** 2) & (s2 < edges[-1] ** 2)
    want = np.histogram2d(np.sqrt(s2[ok]), mu[ok], bins=[edges, np.linspace(0, mu_max, nmu + 1)])[0] * 2
    a = H.oracle_theory("DDsmu_mocks", ra, dec, d, edges, mu_max=mu_max, nmu_bins=nmu, periodic=False)
    assert np.abs(a["npairs"].astype(np.int64) - want.astype(np.int64)).sum() <= 4


@pytest.mark.skipif(H.load_ref() is None or not hasattr(H.load_ref(), "countpairs_mocks"),
                    reason="oracle/_ref was not prebuilt with the mocks statistics")
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("refine", [(2, 2, 1), (1, 1, 1), (3, 2, 2)])
def test_oracle_mocks_vs_live_reference(dtype, refine):
    from corrfunc_b200 import _capi as capi

    ref = H.load_ref()
    ra, dec, d, w = H.mock_points(31, 20000, dtype)
    edges = np.logspace(np.log10(1.0), np.log10(25.0), 8)
    custom = refine != (2, 2, 1)
    o = capi.default_options(dtype, isa=H.ref_isa(), is_comoving_dist=True, bin_refine_factors=refine, custom_refine=custom)
    r = capi.call_DDrppi_mocks(ref, 1, 1, 4, 20.0, edges, ra, dec, d, options=o)
    a = H.oracle_theory("DDrppi_mocks", ra, dec, d, edges, pimax=20.0, periodic=False, refine=refine, custom_refine=custom)
    assert np.array_equal(a["npairs"], r["npairs"])
    o = capi.default_options(dtype, isa=H.ref_isa(), is_comoving_dist=True, bin_refine_factors=refine, custom_refine=custom)
    r = capi.call_DDsmu_mocks(ref, 1, 1, 4, 1.0, 7, edges, ra, dec, d, options=o)
    a = H.oracle_theory("DDsmu_mocks", ra, dec, d, edges, mu_max=1.0, nmu_bins=7, periodic=False, refine=refine, custom_refine=custom)
    assert np.array_equal(a["npairs"], r["npairs"])


@pytest.mark.parametrize("stat", ["DD", "xi"])
def test_oracle_literal_mode_pins_the_float_z_window_of_DD_and_xi(stat):
    """Float32, coordinates 50x rmax, ~2 particles per cell: the reference never visits one pair that lies inside
    the last bin (its per-primary window |dz| < max_dz closes a rounding error early).  LITERAL mode reproduces the
    reference bit for bit; the default mode (every pair, the reference's own r2 arithmetic) counts that pair --
    the same +2 in the last bin, and nothing else, that separates the GPU from the reference."""
    ref = np.load(os.path.join(H.GOLDEN, "ref_float_window.npz"))[stat].astype(np.int64)
    x, y, z, L, edges = H.float_window_case(stat)
    lib = H.load_oracle()
    lib.oracle_set_literal_kernels(1)
    try:
        lit = H.oracle_theory(stat, x, y, z, edges, periodic=True, boxsize=L)["npairs"].astype(np.int64)
    finally:
        lib.oracle_set_literal_kernels(0)
    assert np.array_equal(lit, ref)
    dflt = H.oracle_theory(stat, x, y, z, edges, periodic=True, boxsize=L)["npairs"].astype(np.int64)
    assert np.array_equal((dflt - ref)[:-1], np.zeros(ref.size - 1, dtype=np.int64)) and (dflt - ref)[-1] == 2


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("stat,periodic", [("DD", True), ("DD", False), ("xi", True)])
def test_oracle_literal_equals_default_for_DD_xi_on_small_inputs(stat, periodic, dtype):
    L, N = 420.0, 60000
    x, y, z, w = H.box_points(6, N, L, dtype)
    edges = np.logspace(np.log10(0.1), np.log10(25.0), 15)
    kw = dict(periodic=periodic, boxsize=L, w1=w, weight_type="pair_product", need_avg=True)
    a = H.oracle_theory(stat, x, y, z, edges, **kw)
    lib = H.load_oracle()
    lib.oracle_set_literal_kernels(1)
    try:
        b = H.oracle_theory(stat, x, y, z, edges, **kw)
    finally:
        lib.oracle_set_literal_kernels(0)
    assert np.array_equal(a["npairs"], b["npairs"])
    assert np.allclose(a["ravg"], b["ravg"], rtol=1e-12) and np.allclose(a["weightavg"], b["weightavg"], rtol=1e-12)


def _cz_to_comoving(cz, cosmology):
    import ctypes as C

    from corrfunc_b200 import _lib

    lib = _lib.load()
    lib.corrfunc_b200_cz_to_comoving.restype = C.c_int
    out = np.zeros_like(cz)
    st = lib.corrfunc_b200_cz_to_comoving(C.c_int(cz.itemsize), C.c_int64(cz.size), cz.ctypes.data_as(C.c_void_p),
                                          C.c_int(cosmology), out.ctypes.data_as(C.c_void_p))
    if st != 0:
        raise RuntimeError("corrfunc_b200_cz_to_comoving failed")
    return out


@pytest.mark.parametrize("cosmology", [1, 2])
def test_redshift_distance_table_is_the_references(cosmology):
    """corrfunc_b200_cosmo_dist_table (host code, no GPU) against a committed sample of the table the reference's
    utils/set_cosmo_dist.c produces -- every 97th of its 3499 entries up to z = 0.35, bit for bit."""
    import ctypes as C

    from corrfunc_b200 import _lib

    g = np.load(os.path.join(H.GOLDEN, "ref_mocks_float64.npz"))
    lib = _lib.load()
    lib.corrfunc_b200_cosmo_dist_table.restype = C.c_int
    zc, dc = np.zeros(10000), np.zeros(10000)
    n = lib.corrfunc_b200_cosmo_dist_table(C.c_double(0.35), C.c_int(10000), zc.ctypes.data_as(C.c_void_p),
                                           dc.ctypes.data_as(C.c_void_p), C.c_int(cosmology))
    assert n == int(g["table%d_n" % cosmology])
    assert np.array_equal(zc[:n:97], g["table%d_zc" % cosmology]) and np.array_equal(dc[:n:97], g["table%d_dc" % cosmology])
    assert lib.corrfunc_b200_cosmo_dist_table(C.c_double(0.35), C.c_int(10000), zc.ctypes.data_as(C.c_void_p),
                                              dc.ctypes.data_as(C.c_void_p), C.c_int(3)) == -1  # unknown cosmology


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_cz_input_matches_committed_reference_outputs(dtype):
    """is_comoving_dist = 0: the library's host-side cz -> distance conversion (table + linear interpolation), fed to
    the oracle, reproduces what the reference returns for the cz input itself."""
    g = np.load(os.path.join(H.GOLDEN, "ref_mocks_%s.npz" % np.dtype(dtype).name))
    ra, dec, d, _ = H.mock_points(int(g["seed"]), int(g["N1"]), dtype)
    ra2, dec2, d2, _ = H.mock_points(int(g["seed"]) + 1, int(g["N2"]), dtype)
    cz, cz2 = (d * dtype(60.0)).astype(dtype), (d2 * dtype(60.0)).astype(dtype)
    # the table's reach is set by the larger of the two sets; its entries are the same either way
    D, D2 = _cz_to_comoving(cz, 2), _cz_to_comoving(cz2, 2)
    a = H.oracle_theory("DDrppi_mocks", ra, dec, D, g["edges"], pimax=float(g["pimax"]), autocorr=False, X2=ra2, Y2=dec2,
                        Z2=D2, periodic=False)
    assert np.array_equal(a["npairs"], g["DDrppi_mocks_cz_cross__npairs"])
    a = H.oracle_theory("DDsmu_mocks", ra, dec, D, g["edges"], mu_max=float(g["mu_max"]), nmu_bins=int(g["nmu"]), periodic=False)
    assert np.array_equal(a["npairs"], g["DDsmu_mocks_cz_auto__npairs"])
    with pytest.raises(RuntimeError):  # z < 1e-4: below the table, where GSL would abort the reference
        _cz_to_comoving(np.full(4, 20.0, dtype=dtype), 1)


def test_oracle_matches_reference_golden_DDrppi_mocks():
    """The reference's own known-answer test for DDrppi_mocks (Corrfunc/tests/test_mocks.py:15-34): autocorrelation of
    the Mr19 mock from RA, DEC and **cz** (cosmology 1, pimax 40, PAIR_PRODUCT weights, rpavg) vs
    mocks/tests/Mr19_mock.DD -- a file written upstream by a build with the real GSL.  All 560 npairs exact and the
    averages to the file's print precision: this pins the whole cz path (distance table, the restated GSL
    interpolation, the Cartesian conversion, the pair-midpoint arithmetic)."""
    ra, dec, cz, w = H.load_mr19_mock_cz()
    bins = H.load_bins_file("mocks_bins.txt")
    gold = H.load_ddrppi_mocks_golden()
    D = _cz_to_comoving(cz, 1)
    a = H.oracle_theory("DDrppi_mocks", ra, dec, D, bins, pimax=40.0, w1=w, weight_type="pair_product", need_avg=True,
                        periodic=False)
    assert np.array_equal(a["npairs"].ravel(), gold["npairs"])
    assert np.allclose(a["ravg"].ravel(), gold["ravg"], atol=1e-8, rtol=1e-6)  # common.py:83-105 tolerances
    assert np.allclose(a["weightavg"].ravel(), gold["weightavg"], atol=1e-8, rtol=1e-6)


# ---- counts-in-spheres on survey catalogues: mocks/vpf_mocks ------------------------------------

def test_oracle_matches_reference_golden_vpf_mocks():
    """The reference's own known-answer test for vpf_mocks (Corrfunc/tests/test_mocks.py:82-110): 10 000 spheres from
    its centres file on the Mr19 mock (cz input, cosmology 1), radii 1..10, p0..p5, vs mocks/tests/Mr19_mock_vpf."""
    ra, dec, cz, _ = H.load_mr19_mock_cz()
    cen = np.loadtxt(H.VPF_CENTERS)
    pN, _ = H.oracle_vpf_mocks(ra, dec, H.cz_to_comoving(cz, 1), cen[:, 0], cen[:, 1], cen[:, 2], 10.0, 10, 6)
    assert np.allclose(pN, H.load_vpf_golden(), atol=1e-9, rtol=1e-6)  # common.py:107-116 tolerances


@pytest.mark.skipif(H.load_ref() is None or not hasattr(H.load_ref(), "countspheres_mocks"),
                    reason="oracle/_ref was not prebuilt with vpf_mocks")
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_oracle_vpf_mocks_vs_live_reference_with_randoms(dtype, tmp_path):
    """No centres file: the reference places the spheres on the randoms that have enough neighbours and writes the
    file; harness.vpf_centres_from_randoms restates that choice and the oracle the counting."""
    from corrfunc_b200 import _capi as capi

    ra, dec, d, _ = H.mock_points(41, 20000, dtype)
    rra, rdec, rd, _ = H.mock_points(42, 3000, dtype)
    cfile = str(tmp_path / "centres.txt")
    o = capi.default_options(dtype, isa=H.ref_isa(), bin_refine_factors=(1, 1, 1), is_comoving_dist=True)
    nc = 150
    r = capi.call_vpf_mocks(H.load_ref(), 12.0, 6, nc, 4, 2, cfile, 1, ra, dec, d, RAND_RA=rra, RAND_DEC=rdec, RAND_CZ=rd,
                            options=o)
    rcube = dtype(max(d.max(), rd.max())) + dtype(1.0)
    xc, yc, zc = H.vpf_centres_from_randoms(rra, rdec, rd, rcube, 12.0, 2, nc)
    written = np.loadtxt(cfile)
    assert xc.size == nc and written.shape == (nc, 4)
    assert np.allclose(written[:, 0], xc, atol=1e-4) and np.allclose(written[:, 2], zc, atol=1e-4) and np.all(written[:, 3] == 12.0)
    pN, rc = H.oracle_vpf_mocks(ra, dec, d, xc, yc, zc, 12.0, 6, 4, dmax_randoms=rd.max())
    assert rc == float(rcube)
    assert np.allclose(pN, r["pN"], atol=1e-6 if dtype == np.float32 else 1e-12)


@pytest.mark.skipif(H.load_ref() is None or not hasattr(H.load_ref(), "countpairs_mocks"), reason="oracle/_ref not prebuilt")
@pytest.mark.parametrize("part,seed", [("box", 11), ("sky", 12)])
def test_oracle_fuzz_against_live_reference(part, seed):
    """A short seeded run of tools/fuzz_oracle_vs_reference.py (random options, all statistics): no mismatch."""
    import subprocess
    import sys

    out = subprocess.run([sys.executable, os.path.join(H.ROOT, "tools", "fuzz_oracle_vs_reference.py"), part, str(seed), "25"],
                         capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr[-1500:]
    assert "mismatches 0" in out.stdout and "ran 25" in out.stdout, out.stdout[-1500:]
