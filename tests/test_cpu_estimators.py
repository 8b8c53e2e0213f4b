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
