#!/usr/bin/env python
"""Generates tests/golden/estimators.npz from the REFERENCE's own Python post-processing functions
(/root/reference/Corrfunc/utils.py: convert_3d_counts_to_cf :27-165, convert_rp_pi_counts_to_wp :167-322).
The reference module needs `future` and `wurlitzer`, which are absent here: they are stubbed (neither is used by
the two functions).  Run in the build container only; the .npz travels with the repo."""
import importlib.util
import os
import sys
import types

import numpy as np

fu = types.ModuleType("future")
fuu = types.ModuleType("future.utils")
fuu.bytes_to_native_str = lambda b: b.decode() if isinstance(b, bytes) else b
fu.utils = fuu
sys.modules.setdefault("future", fu)
sys.modules.setdefault("future.utils", fuu)
sys.modules.setdefault("wurlitzer", types.ModuleType("wurlitzer"))
spec = importlib.util.spec_from_file_location("ref_utils", "/root/reference/Corrfunc/utils.py")
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)

rng = np.random.default_rng(2024)
out = {}
# 3-D counts -> xi (Landy-Szalay), including an empty RR bin (-> NaN)
nb = 12
ND1, ND2, NR1, NR2 = 100000, 80000, 300000, 250000
rr = rng.integers(1000, 10 ** 7, nb).astype(np.uint64)
rr[3] = 0
dd = rng.integers(0, 10 ** 6, nb).astype(np.uint64)
d1r2 = rng.integers(0, 3 * 10 ** 6, nb).astype(np.uint64)
d2r1 = rng.integers(0, 3 * 10 ** 6, nb).astype(np.uint64)
out.update(cf_N=np.array([ND1, ND2, NR1, NR2]), cf_dd=dd, cf_d1r2=d1r2, cf_d2r1=d2r1, cf_rr=rr,
           cf_out=ref.convert_3d_counts_to_cf(ND1, ND2, NR1, NR2, dd, d1r2, d2r1, rr))
# rp-pi counts -> wp
nrp, pimax, dpi = 7, 20.0, 1.0
npi = int(pimax)
n2 = nrp * npi
rr2 = rng.integers(1000, 10 ** 7, n2).astype(np.uint64)
dd2 = rng.integers(0, 10 ** 6, n2).astype(np.uint64)
dr2 = rng.integers(0, 3 * 10 ** 6, n2).astype(np.uint64)
out.update(wp_N=np.array([ND1, ND1, NR1, NR1]), wp_dd=dd2, wp_dr=dr2, wp_rr=rr2, wp_nrp=nrp, wp_pimax=pimax, wp_dpi=dpi,
           wp_out=ref.convert_rp_pi_counts_to_wp(ND1, ND1, NR1, NR1, dd2, dr2, dr2, rr2, nrp, pimax, dpi=dpi))
# the smaller helpers of Corrfunc/utils.py (gridlink_sphere :599-863, compute_nbins :521-596)
SPHERE_CASES = [dict(thetamax=10.0), dict(thetamax=3.0, link_in_ra=False),
                dict(thetamax=2.5, ra_limits=[20.0, 200.0], dec_limits=[-30.0, 65.0], ra_refine_factor=2, dec_refine_factor=3),
                dict(thetamax=0.4, max_ra_cells=37, max_dec_cells=50), dict(thetamax=0.2, dec_limits=[-1.5, 1.5], input_in_degrees=False)]
for i, kw in enumerate(SPHERE_CASES):
    if kw.get("link_in_ra", True):
        grid, nra = ref.gridlink_sphere(return_num_ra_cells=True, **kw)
        out["sphere%d_nra" % i] = nra
    else:
        grid = ref.gridlink_sphere(**kw)
    out["sphere%d_dec" % i], out["sphere%d_ra" % i] = grid["dec_limit"], grid["ra_limit"]
out["nbins_cases"] = np.array([[180.0, 10.0, 1, 0], [180.0, 10.0, 2, 20], [0.5, 10.0, 3, 0], [359.9, 0.7, 2, 100]])
out["nbins_out"] = np.array([ref.compute_nbins(a, b, refine_factor=int(c), max_nbins=int(d) or None) for a, b, c, d in out["nbins_cases"]])
np.savez(os.path.join(os.path.dirname(os.path.abspath(__file__)), "estimators.npz"), **out)
print("wrote estimators.npz", {k: np.asarray(v).shape for k, v in out.items()})
