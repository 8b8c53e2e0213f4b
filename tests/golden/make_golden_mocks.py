#!/usr/bin/env python
"""Golden outputs of the UNMODIFIED reference (oracle/_ref, AVX-512F kernels; GSL stand-in, comoving distances) for
mocks/DDrppi_mocks and mocks/DDsmu_mocks on seeded synthetic survey wedges (tests/harness.py:mock_points).
  python tests/golden/make_golden_mocks.py        -> tests/golden/ref_mocks_{float64,float32}.npz
Run in the build container only (oracle/_ref is built from /root/reference)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import harness as H  # noqa: E402
from corrfunc_b200 import _capi  # noqa: E402

ref = H.load_ref()
assert ref is not None and hasattr(ref, "countpairs_mocks"), "build oracle/_ref first (bash oracle/build_ref.sh)"
N1, N2, SEED, PIMAX, MU_MAX, NMU = 60000, 40000, 11, 40.0, 0.9, 10
edges = np.logspace(np.log10(0.5), np.log10(30.0), 13)
for dtype in (np.float64, np.float32):
    ra, dec, d, w = H.mock_points(SEED, N1, dtype)
    ra2, dec2, d2, w2 = H.mock_points(SEED + 1, N2, dtype)
    out = dict(seed=SEED, N1=N1, N2=N2, edges=edges, pimax=PIMAX, mu_max=MU_MAX, nmu=NMU)
    for autocorr in (1, 0):
        tag = "auto" if autocorr else "cross"
        kw = dict(w1=w, weight_type="pair_product")
        if not autocorr:
            kw.update(RA2=ra2, DEC2=dec2, CZ2=d2, w2=w2)
        o = _capi.default_options(dtype, need_avg_sep=True, isa=H.ref_isa(), is_comoving_dist=True)
        r = _capi.call_DDrppi_mocks(ref, autocorr, 1, os.cpu_count(), PIMAX, edges, ra, dec, d, options=o, **kw)
        for k in ("npairs", "ravg", "weightavg"):
            out["DDrppi_mocks_%s__%s" % (tag, k)] = r[k]
        o = _capi.default_options(dtype, need_avg_sep=True, isa=H.ref_isa(), is_comoving_dist=True)
        r = _capi.call_DDsmu_mocks(ref, autocorr, 1, os.cpu_count(), MU_MAX, NMU, edges, ra, dec, d, options=o, **kw)
        for k in ("npairs", "ravg", "weightavg"):
            out["DDsmu_mocks_%s__%s" % (tag, k)] = r[k]
        print(np.dtype(dtype).name, tag, int(out["DDrppi_mocks_%s__npairs" % tag].sum()), int(out["DDsmu_mocks_%s__npairs" % tag].sum()))
    # cz input (is_comoving_dist = 0): the reference's own distance table (utils/set_cosmo_dist.c, compiled unmodified)
    # + GSL's linear interpolation as restated in oracle/gsl_shim; cz = 60 * (the distances above), cosmology 2
    cz, cz2 = (d * dtype(60.0)).astype(dtype), (d2 * dtype(60.0)).astype(dtype)
    o = _capi.default_options(dtype, need_avg_sep=True, isa=H.ref_isa(), is_comoving_dist=False)
    r = _capi.call_DDrppi_mocks(ref, 0, 2, os.cpu_count(), PIMAX, edges, ra, dec, cz, RA2=ra2, DEC2=dec2, CZ2=cz2, options=o)
    out["DDrppi_mocks_cz_cross__npairs"], out["DDrppi_mocks_cz_cross__ravg"] = r["npairs"], r["ravg"]
    o = _capi.default_options(dtype, need_avg_sep=True, isa=H.ref_isa(), is_comoving_dist=False)
    r = _capi.call_DDsmu_mocks(ref, 1, 2, os.cpu_count(), MU_MAX, NMU, edges, ra, dec, cz, options=o)
    out["DDsmu_mocks_cz_auto__npairs"], out["DDsmu_mocks_cz_auto__ravg"] = r["npairs"], r["ravg"]
    print(np.dtype(dtype).name, "cz", int(out["DDrppi_mocks_cz_cross__npairs"].sum()), int(out["DDsmu_mocks_cz_auto__npairs"].sum()))
    # a sample of the reference's redshift -> distance table itself
    import ctypes as C
    ref.set_cosmo_dist.restype = C.c_int
    for cosmo in (1, 2):
        zc, dc = np.zeros(10000), np.zeros(10000)
        n = ref.set_cosmo_dist(C.c_double(0.35), C.c_int(10000), zc.ctypes.data_as(C.c_void_p), dc.ctypes.data_as(C.c_void_p), C.c_int(cosmo))
        out["table%d_n" % cosmo], out["table%d_zc" % cosmo], out["table%d_dc" % cosmo] = n, zc[:n:97].copy(), dc[:n:97].copy()
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ref_mocks_%s.npz" % np.dtype(dtype).name), **out)
