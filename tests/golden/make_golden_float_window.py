#!/usr/bin/env python
"""Outputs of the UNMODIFIED reference (oracle/_ref, AVX-512F kernels) on the two float32 inputs of
tests/harness.py:float_window_case, where its z-window pruning (theory/DD/countpairs_kernels.c.src:104-137,197-203;
theory/xi/xi_kernels.c.src likewise) never visits one pair that lies inside the last bin.  The same effect removes
2 of 8.8e12 pairs at BASELINE config 5 (ref_fullsize_c5.npz vs the GPU: last bin only).
  python tests/golden/make_golden_float_window.py     -> tests/golden/ref_float_window.npz"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import harness as H  # noqa: E402
from corrfunc_b200 import _capi  # noqa: E402

ref = H.load_ref()
assert ref is not None, "build oracle/_ref first"
out = {}
for stat in H.FLOAT_WINDOW_CASES:
    x, y, z, L, edges = H.float_window_case(stat)
    o = _capi.default_options(np.float32, periodic=True, boxsize=L, isa=H.ref_isa())
    r = (_capi.call_xi(ref, L, os.cpu_count(), edges, x, y, z, options=o) if stat == "xi"
         else _capi.call_DD(ref, 1, os.cpu_count(), edges, x, y, z, options=o))
    out[stat] = np.asarray(r["npairs"], dtype=np.uint64)
    print(stat, out[stat])
np.savez_compressed(os.path.join(H.GOLDEN, "ref_float_window.npz"), **out)
