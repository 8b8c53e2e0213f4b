"""Post-processing estimators against golden vectors produced by the reference's own Python functions
(tests/golden/make_golden_estimators.py imported /root/reference/Corrfunc/utils.py)."""
import os

import numpy as np
import pytest

from corrfunc_b200.utils import convert_3d_counts_to_cf, convert_rp_pi_counts_to_wp

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "estimators.npz"))


def test_convert_3d_counts_to_cf_matches_reference():
    N = G["cf_N"]
    got = convert_3d_counts_to_cf(N[0], N[1], N[2], N[3], G["cf_dd"], G["cf_d1r2"], G["cf_d2r1"], G["cf_rr"])
    want = G["cf_out"]
    assert np.array_equal(np.isnan(got), np.isnan(want)) and np.isnan(want).sum() == 1
    ok = ~np.isnan(want)
    assert np.allclose(got[ok], want[ok], rtol=1e-14, atol=0)


def test_structured_counts_are_accepted():
    N = G["cf_N"]
    st = [np.zeros(len(G["cf_dd"]), dtype=[("rmin", "f8"), ("npairs", "u8")]) for _ in range(4)]
    for s, k in zip(st, ("cf_dd", "cf_d1r2", "cf_d2r1", "cf_rr")):
        s["npairs"] = G[k]
    got = convert_3d_counts_to_cf(N[0], N[1], N[2], N[3], *st)
    assert np.allclose(got, G["cf_out"], rtol=1e-14, equal_nan=True)


def test_convert_rp_pi_counts_to_wp_matches_reference():
    N = G["wp_N"]
    got = convert_rp_pi_counts_to_wp(N[0], N[1], N[2], N[3], G["wp_dd"], G["wp_dr"], G["wp_dr"], G["wp_rr"],
                                     int(G["wp_nrp"]), float(G["wp_pimax"]), dpi=float(G["wp_dpi"]))
    assert np.allclose(got, G["wp_out"], rtol=1e-13, atol=0)


def test_estimator_errors():
    a = np.ones(6)
    with pytest.raises(ValueError):
        convert_3d_counts_to_cf(1, 1, 1, 1, a, a, a, a, estimator="Hamilton")
    with pytest.raises(ValueError):
        convert_3d_counts_to_cf(1, 1, 1, 1, a, a[:5], a, a)
    with pytest.raises(ValueError):
        convert_rp_pi_counts_to_wp(1, 1, 1, 1, a, a, a, a, 4, 3.0)  # 6 bins are not a multiple of 4
    with pytest.raises(ValueError):
        convert_rp_pi_counts_to_wp(1, 1, 1, 1, a, a, a, a, 2, 4.0)  # 3 pi bins x dpi 1 != pimax 4
    with pytest.raises(ValueError):
        convert_rp_pi_counts_to_wp(1, 1, 1, 1, a, a, a, a, 2, 3.0, dpi=0.0)


# ---- the smaller helpers of Corrfunc/utils.py ----------------------------------------------------------------------

SPHERE_CASES = [dict(thetamax=10.0), dict(thetamax=3.0, link_in_ra=False),
                dict(thetamax=2.5, ra_limits=[20.0, 200.0], dec_limits=[-30.0, 65.0], ra_refine_factor=2, dec_refine_factor=3),
                dict(thetamax=0.4, max_ra_cells=37, max_dec_cells=50), dict(thetamax=0.2, dec_limits=[-1.5, 1.5], input_in_degrees=False)]


@pytest.mark.parametrize("i", range(len(SPHERE_CASES)))
def test_gridlink_sphere_matches_reference(i):
    from corrfunc_b200.utils import gridlink_sphere

    kw = SPHERE_CASES[i]
    if kw.get("link_in_ra", True):
        grid, nra = gridlink_sphere(return_num_ra_cells=True, **kw)
        assert np.array_equal(nra, G["sphere%d_nra" % i])
    else:
        grid = gridlink_sphere(**kw)
    assert grid.dtype.names == ("dec_limit", "ra_limit")
    assert np.array_equal(grid["dec_limit"], G["sphere%d_dec" % i])
    assert np.array_equal(grid["ra_limit"], G["sphere%d_ra" % i])


def test_compute_nbins_and_small_helpers(tmp_path):
    from corrfunc_b200.utils import (compute_nbins, convert_to_native_endian, fix_cz, fix_ra_dec, is_native_endian,
                                     return_file_with_rbins, sys_pipes)

    got = [compute_nbins(a, b, refine_factor=int(c), max_nbins=int(d) or None) for a, b, c, d in G["nbins_cases"]]
    assert np.array_equal(got, G["nbins_out"])
    with pytest.raises(ValueError):
        compute_nbins(-1.0, 1.0)
    with pytest.raises(ValueError):
        compute_nbins(1.0, 1.0, refine_factor=0)
    # bins: an existing file passes through, an array becomes a temporary "low high" file the C parser reads back exactly
    edges = np.logspace(-1, 1.3, 9)
    name, delete = return_file_with_rbins(edges[::-1])
    assert delete
    rows = np.loadtxt(name)
    os.remove(name)
    assert np.array_equal(rows[:, 0], edges[:-1]) and np.array_equal(rows[:, 1], edges[1:])
    p = str(tmp_path / "bins")
    open(p, "w").write("0.1 1.0\n")
    assert return_file_with_rbins(p) == (p, False)
    with pytest.raises(IOError):
        return_file_with_rbins(str(tmp_path / "missing"))
    # redshift / angle fixes act in place and keep the dtype
    z = np.array([0.01, 0.2], dtype=np.float32)
    out = fix_cz(z)
    assert out.dtype == np.float32 and np.allclose(out, [2998.0, 59960.0]) and np.allclose(z, out)
    cz = np.array([3000.0, 60000.0])
    assert np.array_equal(fix_cz(cz), [3000.0, 60000.0])
    ra, dec = np.array([-170.0, 10.0]), np.array([5.0, 170.0])
    ra2, dec2 = fix_ra_dec(ra, dec)
    assert np.array_equal(ra2, [10.0, 190.0]) and np.array_equal(dec2, [-85.0, 80.0])
    with pytest.raises(TypeError):
        fix_cz([0.1, 0.2])
    # byte order
    swapped = np.arange(10, dtype=np.dtype("i4").newbyteorder("S"))
    assert not is_native_endian(swapped) and is_native_endian(np.arange(3)) and is_native_endian(None)
    native = convert_to_native_endian(swapped)
    assert is_native_endian(native) and np.array_equal(native, np.arange(10))
    a = np.arange(4.0)
    assert convert_to_native_endian(a) is a and convert_to_native_endian(None) is None
    with pytest.warns(UserWarning):
        convert_to_native_endian(swapped, warn=True)
    with sys_pipes():
        pass
