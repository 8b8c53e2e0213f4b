#!/usr/bin/env python
"""Full-size goldens: the UNMODIFIED reference (oracle/_ref, AVX-512F kernels) run on the exact synthetic inputs of the
BASELINE configs as bench.py generates them (same seeds, same dtype), npairs saved to tests/golden/ref_fullsize_<cfg>.npz.
  python tests/golden/make_golden_fullsize.py c1 c2 c2wp32 c2rppi c2rppi32 c3 c4      (about two minutes on 8 cores)
  python tests/golden/make_golden_fullsize.py c5sd10M                                 (config 5 at the same density, 10 M points)
  python tests/golden/make_golden_fullsize.py c5                                      (100 M points: ~40 minutes)
Run in the build container only (needs /root/reference to build oracle/_ref)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
import harness as H  # noqa: E402
from corrfunc_b200 import _capi  # noqa: E402

ref = H.load_ref()
assert ref is not None, "build oracle/_ref first (python -c 'import __graft_entry__ as g; g.build()')"
nthreads = int(os.environ.get("GOLDEN_THREADS", os.cpu_count()))
for name in sys.argv[1:]:
    cfg = bench.config_by_name(name)
    dtype = np.float32 if cfg["dtype"] == "f32" else np.float64
    bins = bench.make_bins(cfg["bins"])
    pts = bench.gen_points(cfg, cfg["N"], dtype)
    o = _capi.default_options(dtype, periodic=True, need_avg_sep=bool(cfg.get("avg")), isa=H.ref_isa(),
                              boxsize=cfg["L"] if cfg["L"] > 0 else None)
    w = pts.get("w")
    wt = "pair_product" if cfg.get("weights") else None
    t0 = time.time()
    st = cfg["stat"]
    if st == "xi":
        r = _capi.call_xi(ref, cfg["L"], nthreads, bins, pts["x"], pts["y"], pts["z"], options=o)
    elif st == "DD":
        r = _capi.call_DD(ref, 1, nthreads, bins, pts["x"], pts["y"], pts["z"], options=o)
    elif st == "wp":
        r = _capi.call_wp(ref, cfg["L"], nthreads, cfg["pimax"], bins, pts["x"], pts["y"], pts["z"], options=o)
    elif st == "DDrppi":
        r = _capi.call_DDrppi(ref, 1, nthreads, cfg["pimax"], bins, pts["x"], pts["y"], pts["z"], options=o)
    elif st == "DDsmu":
        r = _capi.call_DDsmu(ref, 1, nthreads, bins, cfg["mu_max"], cfg["nmu"], pts["x"], pts["y"], pts["z"], w1=w,
                             weight_type=wt, options=o)
    elif st == "DDtheta":
        r = _capi.call_DDtheta(ref, 0, nthreads, bins, pts["ra"], pts["dec"], RA2=pts["ra2"], DEC2=pts["dec2"], options=o)
    dt = time.time() - t0
    out = {"npairs": np.asarray(r["npairs"], dtype=np.uint64), "seconds": dt, "nthreads": nthreads}
    for k in ("ravg", "weightavg"):
        if k in r and (cfg.get("avg") or cfg.get("weights")):
            out[k] = np.asarray(r[k], dtype=np.float64)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ref_fullsize_%s.npz" % name), **out)
    print("%s: reference %s on N=%d took %.1f s, sum(npairs)=%d" % (name, st, cfg["N"], dt, int(out["npairs"].sum())), flush=True)
