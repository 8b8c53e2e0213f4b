#!/usr/bin/env python
"""Does the cell sharding itself cost kernel time?  Config 5 (or a same-density subsample) on ONE GPU with the library told
it is rank r of n: the kernel then counts only that rank's cells.  If t(rank r of n) is close to t(all) / n the sharding
pattern is free and the loss seen at 8 GPUs is a difference between devices; if not, it is the access pattern.
  python tools/exp_shard.py [config] [n] [ranks...]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from corrfunc_b200 import _lib  # noqa: E402
from corrfunc_b200.theory import xi  # noqa: E402

cfg = bench.config_by_name(sys.argv[1] if len(sys.argv) > 1 else "c5")
n = int(sys.argv[2]) if len(sys.argv) > 2 else 8
ranks = [int(a) for a in sys.argv[3:]] or [0, n // 2]
pts = bench.gen_points(cfg, cfg["N"], np.float32)
bins = bench.make_bins(cfg["bins"])
x, y, z = pts["x"], pts["y"], pts["z"]


def run(name, reps=2):
    best = 1e30
    for _ in range(reps):
        xi(cfg["L"], 1, bins, x, y, z)
        st = _lib.last_stats()
        best = min(best, st["ms_pairs"])
    print("%-22s kern %9.2f ms  grid %6.2f ms  n_eval %.4e" % (name, best, st["ms_gridlink"], st["n_eval"]), flush=True)
    return best, st["n_eval"]


t_all, e_all = run("all cells")
for r in ranks:
    _lib.set_shard(r, n, lambda a, b, c: None)  # identity "reduction": the partial result is all we want
    t, e = run("rank %d of %d" % (r, n))
    print("   -> %.4f of the work in %.4f of the time: sharding costs %+.2f %%" % (e / e_all, t / t_all, 100.0 * (t / t_all / (e / e_all) - 1.0)))
_lib.set_shard(0, 1, None)
