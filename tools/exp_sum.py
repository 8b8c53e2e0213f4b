#!/usr/bin/env python
"""Where does the per-pair-sum kernel spend its time?  c3-like DDsmu (same density, N points) with the per-pair
outputs switched on one at a time; kernel time from the library's CUDA events.  Run on the GPU box:
  python tools/exp_sum.py [N]      (env knobs: CORRFUNC_B200_SUM_OCC / _SUM_BLOCKS / _SUM_COPIES)"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from corrfunc_b200 import _lib  # noqa: E402
from corrfunc_b200.theory import DD, DDrppi, DDsmu  # noqa: E402

N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 3_000_000
cfg = bench.config_by_name("c3sd%gM" % (N / 1e6))
pts = bench.gen_points(cfg, cfg["N"], np.float64)
bins = bench.make_bins(cfg["bins"])
L = cfg["L"]


def run(name, fn):
    best = 1e30
    for _ in range(3):
        fn()
        st = _lib.last_stats()
        best = min(best, st["ms_pairs"])
    print("%-34s kern %8.2f ms  n_eval %.3e  jobs %.3e  kind %d  fine %s" % (name, best, st["n_eval"], st["n_tilepairs"],
                                                                            st["kernel_kind"], st["fine"]), flush=True)


x, y, z, w = pts["x"], pts["y"], pts["z"], pts["w"]
run("DDsmu count only", lambda: DDsmu(1, 1, bins, 1.0, 20, x, y, z, periodic=True, boxsize=L))
run("DDsmu + savg", lambda: DDsmu(1, 1, bins, 1.0, 20, x, y, z, periodic=True, boxsize=L, output_savg=True))
run("DDsmu + weights", lambda: DDsmu(1, 1, bins, 1.0, 20, x, y, z, weights1=w, weight_type="pair_product", periodic=True, boxsize=L))
run("DDsmu + savg + weights (c3)", lambda: DDsmu(1, 1, bins, 1.0, 20, x, y, z, weights1=w, weight_type="pair_product", periodic=True,
                                                boxsize=L, output_savg=True))
_lib.load().cfb_force_kernel(0)
run("DD (sum kernel) count only", lambda: DD(1, 1, bins, x, y, z, periodic=True, boxsize=L))
run("DD (sum kernel) + ravg", lambda: DD(1, 1, bins, x, y, z, periodic=True, boxsize=L, output_ravg=True))
run("DDrppi pimax=40 count only", lambda: DDrppi(1, 1, 40.0, bins, x, y, z, periodic=True, boxsize=L))
_lib.load().cfb_force_kernel(3)
run("legacy: DDsmu + savg + weights", lambda: DDsmu(1, 1, bins, 1.0, 20, x, y, z, weights1=w, weight_type="pair_product", periodic=True,
                                                   boxsize=L, output_savg=True))
run("legacy: DDsmu count only", lambda: DDsmu(1, 1, bins, 1.0, 20, x, y, z, periodic=True, boxsize=L))
_lib.load().cfb_force_kernel(-1)
