"""Loader for the product library ``csrc/libcorrfunc_b200.so`` (C host layer + CUDA kernels).

There is no CPU fallback: if the library is missing this raises, and every call into it fails with
EXIT_FAILURE when no CUDA device is present.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CORRFUNC_B200_LIBPATH") or os.path.join(_HERE, "csrc", "libcorrfunc_b200.so")
_lib = None
_hook_keepalive = None


class CfbStats(C.Structure):
    _fields_ = [("ms_h2d", C.c_double), ("ms_gridlink", C.c_double), ("ms_pairs", C.c_double),
                ("ms_total_device", C.c_double), ("n_eval", C.c_uint64), ("n_tilepairs", C.c_uint64),
                ("n_cells", C.c_int64), ("n_tiles", C.c_int64), ("fine", C.c_int * 3),
                ("kernel_launches", C.c_int), ("kernel_kind", C.c_int), ("n_analytic", C.c_uint64), ("n_levelpairs", C.c_uint64)]


class Stats(C.Structure):
    _fields_ = [("dev", CfbStats), ("ms_host_total", C.c_double), ("ms_upload", C.c_double),
                ("nmesh", C.c_int * 3), ("refine", C.c_int * 3)]


REDUCE_FN = C.CFUNCTYPE(C.c_int, C.POINTER(C.c_uint64), C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int64,
                        C.c_void_p)


def load():
    """Return the ctypes handle of libcorrfunc_b200.so (built by ``make -C corrfunc_b200/csrc`` or
    ``__graft_entry__.build()``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "corrfunc_b200: %s is missing -- build it with `make -C corrfunc_b200/csrc` "
            "(there is no CPU fallback)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH, mode=os.RTLD_LOCAL)
    lib.corrfunc_b200_last_stats.restype = C.POINTER(Stats)
    lib.corrfunc_b200_version.restype = C.c_char_p
    lib.corrfunc_b200_set_shard.argtypes = [C.c_int, C.c_int]
    lib.corrfunc_b200_set_reduce_hook.argtypes = [REDUCE_FN, C.c_void_p]
    lib.cfb_last_error.restype = C.c_char_p
    lib.cfb_set_target_occupancy.argtypes = [C.c_int]
    lib.cfb_force_kernel.argtypes = [C.c_int]
    lib.cfb_init.restype = C.c_int
    lib.corrfunc_b200_catalog_cache.argtypes = [C.c_int]
    lib.cfb_catalog_cache_hits.restype = C.c_longlong
    lib.cfb_last_device_count.restype = C.c_int
    _lib = lib
    return lib


def last_stats() -> dict:
    s = load().corrfunc_b200_last_stats().contents
    d = s.dev
    return dict(ms_gridlink=d.ms_gridlink, ms_pairs=d.ms_pairs, ms_total_device=d.ms_total_device,
                n_eval=int(d.n_eval), n_tilepairs=int(d.n_tilepairs), n_cells=int(d.n_cells),
                n_tiles=int(d.n_tiles), fine=tuple(d.fine), kernel_launches=int(d.kernel_launches),
                kernel_kind=int(d.kernel_kind), n_analytic=int(d.n_analytic), n_levelpairs=int(d.n_levelpairs), ms_host_total=s.ms_host_total, ms_upload=s.ms_upload,
                nmesh=tuple(s.nmesh), refine=tuple(s.refine))


def set_shard(rank: int, nranks: int, reduce_fn=None):
    """Shard the primary cells over `nranks` processes (one per GPU).  `reduce_fn(npairs, sum_sep,
    sum_w)` receives numpy views of the raw per-bin histograms and must sum them across ranks in
    place (see corrfunc_b200.parallel.enable_distributed)."""
    global _hook_keepalive
    import numpy as np

    lib = load()
    lib.corrfunc_b200_set_shard(int(rank), int(nranks))
    if reduce_fn is None:
        _hook_keepalive = None
        lib.corrfunc_b200_set_reduce_hook(C.cast(None, REDUCE_FN), None)
        return

    def _hook(np_p, ss_p, sw_p, n, _user):
        try:
            a = np.ctypeslib.as_array(np_p, shape=(n,))
            b = np.ctypeslib.as_array(ss_p, shape=(n,))
            c = np.ctypeslib.as_array(sw_p, shape=(n,))
            reduce_fn(a, b, c)
            return 0
        except Exception as exc:  # pragma: no cover - surfaced as EXIT_FAILURE by the C layer
            print("corrfunc_b200 reduce hook failed: %r" % (exc,))
            return 1

    _hook_keepalive = REDUCE_FN(_hook)
    lib.corrfunc_b200_set_reduce_hook(_hook_keepalive, None)
