#!/usr/bin/env python
"""Pairs the reference's float DDrppi kernel drops on BASELINE config 2 (c2rppi32: 1.2 M points, float).

countpairs_rp_pi_avx512_intrinsics takes |dz| BEFORE its "some lane reached pimax -> this is the last chunk" test
(countpairs_rp_pi_kernels.c.src:196-207).  In float a secondary that survived the fast-forward (z1 > zpos - pimax)
can still round to dz == -pimax exactly; the kernel then stops after that 16-lane chunk and every later secondary
of the primary is lost.  The oracle's LITERAL mode follows that control flow and is bit-identical to the
reference's output (ref_fullsize_c2rppi32.npz); its default mode counts every pair that meets the reference's
per-pair conditions, which is also what the GPU computes.  This script stores default - literal.
  python tests/golden/make_golden_rppi32_dropped.py      (~30 s on 8 cores; needs only oracle/, not /root/reference)"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
import harness as H  # noqa: E402

name = "c2rppi32"
out = {}
for literal in (0, 1):
    out[literal] = H.oracle_config(name, literal=literal)["npairs"].astype(np.int64)
ref = np.load(os.path.join(H.GOLDEN, "ref_fullsize_%s.npz" % name))["npairs"].astype(np.int64)
assert np.array_equal(out[1].reshape(ref.shape), ref), "literal oracle != reference golden"
dropped = (out[0] - out[1]).reshape(ref.shape)
assert dropped.min() >= 0
np.savez_compressed(os.path.join(H.GOLDEN, "ref_fullsize_%s_dropped.npz" % name), dropped=dropped.astype(np.uint64))
print("dropped ordered pairs: %d in %d bins, max %d per bin" % (dropped.sum(), np.count_nonzero(dropped), dropped.max()))
