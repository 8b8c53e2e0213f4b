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
np.savez(os.path.join(os.path.dirname(os.path.abspath(__file__)), "estimators.npz"), **out)
print("wrote estimators.npz", {k: np.asarray(v).shape for k, v in out.items()})
