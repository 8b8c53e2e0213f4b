#!/usr/bin/env python
"""Dump the GPU's npairs of BASELINE configs (bench.py inputs) to gpurun_out/gpu_fullsize_<cfg>.npy, for an
offline diff against tests/golden/ref_fullsize_<cfg>.npz.   python tools/dump_fullsize.py c2wp32 c2rppi32"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_gpu_parity as P  # noqa: E402
from corrfunc_b200 import _lib  # noqa: E402

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
lib = _lib.load()
for spec in sys.argv[1:]:
    name, _, kind = spec.partition(":")
    lib.cfb_force_kernel({"": -1, "generic": 0, "fast": 1}[kind])
    r = P._run_config(lib, name)  # bench.config_by_name: c1..c5 or e.g. c5sd10M
    np.save(os.path.join(ROOT, "gpurun_out", "gpu_fullsize_%s.npy" % spec.replace(":", "_")), np.asarray(r["npairs"], dtype=np.uint64))
    print(spec, int(np.asarray(r["npairs"], dtype=np.uint64).sum()))
