"""Configs 1, 3 and 4 on float32 inputs at full size against the reference's float path (goldens:
tests/golden/ref_fullsize_c{1,3,4}f32.npz; the CPU oracle agrees with the reference on all three).  Added after the
round's GPU minutes were spent: this file sorts last so that its first GPU run cannot mask the verified suites."""
import pytest

import test_gpu_parity as P

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["c1f32", "c4f32", "c3f32"])
def test_full_size_float_twin_vs_reference_golden(name):
    P._check_full_size(name)


# ---- cz input of the mocks statistics (is_comoving_dist = 0): host-side conversion, added with the tests above -------

import os  # noqa: E402

import numpy as np  # noqa: E402

import harness as H  # noqa: E402


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_mocks_cz_input_vs_reference_golden(dtype):
    """countpairs_mocks / countpairs_mocks_s_mu with cz (km/s) instead of distances, against the committed outputs of
    the reference for the same cz input (its own distance table + GSL's linear interpolation as restated in
    oracle/gsl_shim).  The conversion is host code and is checked bit for bit on the CPU (tests/test_cpu_oracle.py)."""
    from corrfunc_b200 import _capi, _lib
    from corrfunc_b200.mocks import DDsmu_mocks

    g = np.load(os.path.join(H.GOLDEN, "ref_mocks_%s.npz" % np.dtype(dtype).name))
    ra, dec, d, _ = H.mock_points(int(g["seed"]), int(g["N1"]), dtype)
    ra2, dec2, d2, _ = H.mock_points(int(g["seed"]) + 1, int(g["N2"]), dtype)
    cz, cz2 = (d * dtype(60.0)).astype(dtype), (d2 * dtype(60.0)).astype(dtype)
    o = _capi.default_options(dtype, need_avg_sep=True, is_comoving_dist=False)
    r = _capi.call_DDrppi_mocks(_lib.load(), 0, 2, 1, float(g["pimax"]), g["edges"], ra, dec, cz, RA2=ra2, DEC2=dec2,
                                CZ2=cz2, options=o)
    assert np.array_equal(r["npairs"], g["DDrppi_mocks_cz_cross__npairs"])
    s = DDsmu_mocks(1, 2, 1, float(g["mu_max"]), int(g["nmu"]), g["edges"], ra, dec, cz)  # the wrapper's default: cz
    assert np.array_equal(s["npairs"], g["DDsmu_mocks_cz_auto__npairs"].ravel())
    # redshifts passed where cz is expected (maximum below 10): scaled by the speed of light like the reference does
    z_in = (cz / dtype(299800.0)).astype(dtype)
    s2 = DDsmu_mocks(1, 2, 1, float(g["mu_max"]), int(g["nmu"]), g["edges"], ra, dec, z_in)
    assert abs(int(s2["npairs"].sum()) - int(s["npairs"].sum())) <= 1e-4 * int(s["npairs"].sum())


def test_DDrppi_mocks_reference_golden_file():
    """The reference's own known-answer test (Corrfunc/tests/test_mocks.py:15-34): DDrppi_mocks autocorr of the Mr19
    mock from RA, DEC, cz vs mocks/tests/Mr19_mock.DD, atol 1e-9 / rtol 1e-6 as in common.py:83-105."""
    from corrfunc_b200.mocks import DDrppi_mocks

    ra, dec, cz, w = H.load_mr19_mock_cz()
    bins = H.load_bins_file("mocks_bins.txt")
    gold = H.load_ddrppi_mocks_golden()
    r = DDrppi_mocks(1, 1, 4, 40.0, bins, ra, dec, cz, weights1=w, weight_type="pair_product", output_rpavg=True)
    assert np.array_equal(r["npairs"], gold["npairs"])
    assert np.allclose(r["rpavg"], gold["ravg"], atol=1e-9, rtol=1e-6)
    assert np.allclose(r["weightavg"], gold["weightavg"], atol=1e-9, rtol=1e-6)


# ---- counts-in-spheres (mocks/vpf_mocks), added with the tests above ------------------------------------------------

def test_vpf_mocks_reference_golden_file():
    """The reference's own known-answer test (Corrfunc/tests/test_mocks.py:82-110): 10 000 spheres from its centres
    file on the Mr19 mock, radii 1..10, p0..p5 vs mocks/tests/Mr19_mock_vpf (atol 1e-9 / rtol 1e-6, common.py:107-116)."""
    from corrfunc_b200.mocks import vpf_mocks

    ra, dec, cz, _ = H.load_mr19_mock_cz()
    r = vpf_mocks(10.0, 10, 10000, 6, 1, H.VPF_CENTERS, 1, ra, dec, cz, ra, dec, cz)
    assert r.dtype.names == ("rmax", "pN") and np.allclose(r["rmax"], np.arange(1, 11))
    assert np.allclose(r["pN"], H.load_vpf_golden(), atol=1e-9, rtol=1e-6)
    # the same through the float path, against the oracle in float
    raf, decf, czf = (a.astype(np.float32) for a in (ra, dec, cz))
    cen = np.loadtxt(H.VPF_CENTERS).astype(np.float32)
    want, _ = H.oracle_vpf_mocks(raf, decf, H.cz_to_comoving(czf, 1), cen[:, 0], cen[:, 1], cen[:, 2], 10.0, 10, 6)
    rf = vpf_mocks(10.0, 10, 10000, 6, 1, H.VPF_CENTERS, 1, raf, decf, czf, raf, decf, czf)
    assert np.array_equal(rf["pN"], want)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_vpf_mocks_centres_from_randoms(dtype, tmp_path):
    """Without a usable centres file the spheres go on the first randoms with enough neighbours and the file is
    rewritten (countspheres_mocks_impl.c.src:478-562): centres, file and pN against the restatement in tests/harness.py
    and the oracle (both checked against the live reference on the CPU)."""
    from corrfunc_b200.mocks import vpf_mocks

    ra, dec, d, _ = H.mock_points(41, 20000, dtype)
    rra, rdec, rd, _ = H.mock_points(42, 3000, dtype)
    cfile = str(tmp_path / "centres.txt")
    nc = 150
    r = vpf_mocks(12.0, 6, nc, 4, 2, cfile, 1, ra, dec, d, rra, rdec, rd, is_comoving_dist=True)
    rcube = dtype(max(d.max(), rd.max())) + dtype(1.0)
    xc, yc, zc = H.vpf_centres_from_randoms(rra, rdec, rd, rcube, 12.0, 2, nc)
    written = np.loadtxt(cfile)
    assert written.shape == (nc, 4) and np.allclose(written[:, 0], xc, atol=1e-4) and np.allclose(written[:, 1], yc, atol=1e-4)
    want, _ = H.oracle_vpf_mocks(ra, dec, d, xc, yc, zc, 12.0, 6, 4, dmax_randoms=rd.max())
    assert np.array_equal(r["pN"], want)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("periodic", [True, False])
def test_theory_vpf_vs_oracle(dtype, periodic):
    """countspheres / Corrfunc.theory.vpf: centres from the MT19937 stream (restated in tests/harness.py from the
    library's own stream, which equals numpy's), counts against the oracle -- both of which agree with the unmodified
    reference on the CPU (tests/test_cpu_host_layer.py)."""
    from corrfunc_b200.theory import vpf

    L, N = 300.0, 40000
    x, y, z, _ = H.box_points(9, N, L, dtype)
    rmax, nbin, nc, num_pN, seed = 12.0, 6, 800, 5, 77
    r = vpf(rmax, nbin, nc, num_pN, seed, x, y, z, periodic=periodic, boxsize=L if periodic else None)
    xc, yc, zc, wrap = H.vpf_theory_centres(x, y, z, rmax, nc, seed, periodic, L)
    a = H.oracle_vpf_theory(x, y, z, xc, yc, zc, periodic, wrap, rmax, nbin, num_pN)
    assert r.dtype.names == ("rmax", "pN") and np.allclose(r["rmax"], 2.0 * np.arange(1, 7))
    assert np.array_equal(r["pN"], a)
    # a box only three cells wide per axis, spheres reaching through the periodic faces
    if periodic:
        xs, ys, zs, _ = H.box_points(10, 3000, 40.0, dtype)
        r = vpf(12.0, 4, 300, 4, 5, xs, ys, zs, periodic=True, boxsize=40.0)
        xc, yc, zc, wrap = H.vpf_theory_centres(xs, ys, zs, 12.0, 300, 5, True, 40.0)
        assert np.array_equal(r["pN"], H.oracle_vpf_theory(xs, ys, zs, xc, yc, zc, True, wrap, 12.0, 4, 4))
