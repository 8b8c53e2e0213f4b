"""``Corrfunc.theory`` look-alikes: same positional/keyword signatures, same structured-array results.

Reference wrappers mirrored here: Corrfunc/theory/DD.py:16-290, DDrppi.py:16-345, DDsmu.py:16-370,
wp.py:283-545, xi.py:18-255.  Differences: ``isa``, ``copy_particles`` and ``enable_min_sep_opt`` are
accepted and ignored (results do not depend on them), and the work runs through the C-ABI library
``libcorrfunc_b200.so`` instead of a CPython extension.
"""
from __future__ import annotations

import numpy as np

from .. import _capi, _lib
from ..utils import check_same_dtype, convert_to_native_endian, native_inputs, translate_isa_string_to_enum


def _options(dtype, *, periodic, boxsize, verbose, need_avg, refine, default_refine, max_cells_per_dim,
             copy_particles, enable_min_sep_opt, c_api_timer, isa):
    translate_isa_string_to_enum(isa)
    custom = tuple(int(r) for r in refine) != tuple(default_refine)  # _countpairs.c:1243-1250
    return _capi.default_options(dtype, verbose=verbose, periodic=periodic, need_avg_sep=need_avg, boxsize=boxsize,
                                 bin_refine_factors=refine, max_cells_per_dim=max_cells_per_dim,
                                 copy_particles=copy_particles, enable_min_sep_opt=enable_min_sep_opt,
                                 c_api_timer=c_api_timer, isa=-1, custom_refine=custom)


def _w0(w):
    return None if w is None else np.ascontiguousarray(w[0])


def DD(autocorr, nthreads, binfile, X1, Y1, Z1, weights1=None, periodic=True, boxsize=None, X2=None, Y2=None,
       Z2=None, weights2=None, verbose=False, output_ravg=False, xbin_refine_factor=2, ybin_refine_factor=2,
       zbin_refine_factor=1, max_cells_per_dim=100, copy_particles=True, enable_min_sep_opt=True,
       c_api_timer=False, isa="fastest", weight_type=None):
    """3-D pair counts DD(r).  Returns a structured array (rmin, rmax, ravg, npairs, weightavg)
    [and the C call's wall time when ``c_api_timer``], like ``Corrfunc.theory.DD``."""
    if not autocorr and (X2 is None or Y2 is None or Z2 is None):
        raise ValueError("Must pass valid arrays for X2/Y2/Z2 for computing cross-correlation")
    if periodic and boxsize is None:
        raise ValueError("Must specify a boxsize if periodic=True")
    (X1, Y1, Z1, X2, Y2, Z2), weights1, weights2, dtype = native_inputs((X1, Y1, Z1, X2, Y2, Z2), weights1, weights2, X1, X2,
                                                                        weight_type, autocorr)
    opt = _options(dtype, periodic=periodic, boxsize=boxsize, verbose=verbose, need_avg=output_ravg,
                   refine=(xbin_refine_factor, ybin_refine_factor, zbin_refine_factor), default_refine=(2, 2, 1),
                   max_cells_per_dim=max_cells_per_dim, copy_particles=copy_particles,
                   enable_min_sep_opt=enable_min_sep_opt, c_api_timer=c_api_timer, isa=isa)
    r = _capi.call_DD(_lib.load(), autocorr, nthreads, binfile, X1, Y1, Z1, w1=_w0(weights1), X2=X2, Y2=Y2, Z2=Z2,
                      w2=_w0(weights2), weight_type=weight_type, options=opt, dtype=dtype)
    res = np.zeros(r["npairs"].size, dtype=[("rmin", np.float64), ("rmax", np.float64), ("ravg", np.float64),
                                            ("npairs", np.uint64), ("weightavg", np.float64)])
    res["rmin"], res["rmax"] = r["rupp"][:-1], r["rupp"][1:]
    res["ravg"], res["npairs"], res["weightavg"] = r["ravg"], r["npairs"], r["weightavg"]
    return (res, r["api_time"]) if c_api_timer else res


def DDrppi(autocorr, nthreads, pimax, binfile, X1, Y1, Z1, weights1=None, periodic=True, boxsize=None, X2=None,
           Y2=None, Z2=None, weights2=None, verbose=False, output_rpavg=False, xbin_refine_factor=2,
           ybin_refine_factor=2, zbin_refine_factor=1, max_cells_per_dim=100, copy_particles=True,
           enable_min_sep_opt=True, c_api_timer=False, isa="fastest", weight_type=None):
    """Pair counts DD(rp, pi); rows ordered rp-major with ``int(pimax)`` unit-width pi bins each."""
    if not autocorr and (X2 is None or Y2 is None or Z2 is None):
        raise ValueError("Must pass valid arrays for X2/Y2/Z2 for computing cross-correlation")
    if periodic and boxsize is None:
        raise ValueError("Must specify a boxsize if periodic=True")
    (X1, Y1, Z1, X2, Y2, Z2), weights1, weights2, dtype = native_inputs((X1, Y1, Z1, X2, Y2, Z2), weights1, weights2, X1, X2,
                                                                        weight_type, autocorr)
    opt = _options(dtype, periodic=periodic, boxsize=boxsize, verbose=verbose, need_avg=output_rpavg,
                   refine=(xbin_refine_factor, ybin_refine_factor, zbin_refine_factor), default_refine=(2, 2, 1),
                   max_cells_per_dim=max_cells_per_dim, copy_particles=copy_particles,
                   enable_min_sep_opt=enable_min_sep_opt, c_api_timer=c_api_timer, isa=isa)
    r = _capi.call_DDrppi(_lib.load(), autocorr, nthreads, pimax, binfile, X1, Y1, Z1, w1=_w0(weights1), X2=X2,
                          Y2=Y2, Z2=Z2, w2=_w0(weights2), weight_type=weight_type, options=opt, dtype=dtype)
    nrp, npi = r["npairs"].shape
    res = np.zeros(nrp * npi, dtype=[("rmin", np.float64), ("rmax", np.float64), ("rpavg", np.float64),
                                     ("pimax", np.float64), ("npairs", np.uint64), ("weightavg", np.float64)])
    dpi = r["pimax"] / npi  # rows as built in _countpairs.c:1722-1740
    res["rmin"] = np.repeat(r["rupp"][:-1], npi)
    res["rmax"] = np.repeat(r["rupp"][1:], npi)
    res["pimax"] = np.tile((np.arange(npi) + 1) * dpi, nrp)
    res["rpavg"], res["npairs"], res["weightavg"] = r["ravg"].ravel(), r["npairs"].ravel(), r["weightavg"].ravel()
    return (res, r["api_time"]) if c_api_timer else res


def DDsmu(autocorr, nthreads, binfile, mu_max, nmu_bins, X1, Y1, Z1, weights1=None, periodic=True, boxsize=None,
          X2=None, Y2=None, Z2=None, weights2=None, verbose=False, output_savg=False, fast_divide_and_NR_steps=0,
          xbin_refine_factor=2, ybin_refine_factor=2, zbin_refine_factor=1, max_cells_per_dim=100,
          copy_particles=True, enable_min_sep_opt=True, c_api_timer=False, isa="fastest", weight_type=None):
    """Pair counts DD(s, mu) with ``nmu_bins`` linear mu bins in [0, mu_max).  The GPU path always
    uses a true IEEE divide for mu^2 (``fast_divide_and_NR_steps`` is accepted and treated as 0)."""
    if not autocorr and (X2 is None or Y2 is None or Z2 is None):
        raise ValueError("Must pass valid arrays for X2/Y2/Z2 for computing cross-correlation")
    if periodic and boxsize is None:
        raise ValueError("Must specify a boxsize if periodic=True")
    if mu_max <= 0.0 or mu_max > 1.0:
        raise ValueError("The parameter `mu_max` = {0}, has to be in (0.0, 1.0]".format(mu_max))
    if nmu_bins < 1:
        raise ValueError("Number of mu bins must be at least 1")
    (X1, Y1, Z1, X2, Y2, Z2), weights1, weights2, dtype = native_inputs((X1, Y1, Z1, X2, Y2, Z2), weights1, weights2, X1, X2,
                                                                        weight_type, autocorr)
    opt = _options(dtype, periodic=periodic, boxsize=boxsize, verbose=verbose, need_avg=output_savg,
                   refine=(xbin_refine_factor, ybin_refine_factor, zbin_refine_factor), default_refine=(2, 2, 1),
                   max_cells_per_dim=max_cells_per_dim, copy_particles=copy_particles,
                   enable_min_sep_opt=enable_min_sep_opt, c_api_timer=c_api_timer, isa=isa)
    r = _capi.call_DDsmu(_lib.load(), autocorr, nthreads, binfile, mu_max, nmu_bins, X1, Y1, Z1, w1=_w0(weights1),
                         X2=X2, Y2=Y2, Z2=Z2, w2=_w0(weights2), weight_type=weight_type, options=opt, dtype=dtype)
    ns, nmu = r["npairs"].shape
    res = np.zeros(ns * nmu, dtype=[("smin", np.float64), ("smax", np.float64), ("savg", np.float64),
                                    ("mu_max", np.float64), ("npairs", np.uint64), ("weightavg", np.float64)])
    dmu = r["mu_max"] / nmu
    res["smin"] = np.repeat(r["rupp"][:-1], nmu)
    res["smax"] = np.repeat(r["rupp"][1:], nmu)
    res["mu_max"] = np.tile((np.arange(nmu) + 1) * dmu, ns)
    res["savg"], res["npairs"], res["weightavg"] = r["ravg"].ravel(), r["npairs"].ravel(), r["weightavg"].ravel()
    return (res, r["api_time"]) if c_api_timer else res


def wp(boxsize, pimax, nthreads, binfile, X, Y, Z, weights=None, weight_type=None, verbose=False,
       output_rpavg=False, xbin_refine_factor=2, ybin_refine_factor=2, zbin_refine_factor=1,
       max_cells_per_dim=100, copy_particles=True, enable_min_sep_opt=True, c_api_timer=False,
       c_cell_timer=False, isa="fastest"):
    """Projected correlation function wp(rp) in a periodic cube.  ``c_cell_timer`` has no GPU
    equivalent (there is no per-cell-pair CPU kernel call to time): the third return value is None."""
    (X, Y, Z), weights, _, dtype = native_inputs((X, Y, Z), weights, None, X, None, weight_type, True)
    opt = _options(dtype, periodic=True, boxsize=boxsize, verbose=verbose, need_avg=output_rpavg,
                   refine=(xbin_refine_factor, ybin_refine_factor, zbin_refine_factor), default_refine=(2, 2, 1),
                   max_cells_per_dim=max_cells_per_dim, copy_particles=copy_particles,
                   enable_min_sep_opt=enable_min_sep_opt, c_api_timer=c_api_timer, isa=isa)
    r = _capi.call_wp(_lib.load(), boxsize, nthreads, pimax, binfile, X, Y, Z, w=_w0(weights),
                      weight_type=weight_type, options=opt, dtype=dtype)
    res = np.zeros(r["npairs"].size, dtype=[("rmin", np.float64), ("rmax", np.float64), ("rpavg", np.float64),
                                            ("wp", np.float64), ("npairs", np.uint64), ("weightavg", np.float64)])
    res["rmin"], res["rmax"] = r["rupp"][:-1], r["rupp"][1:]
    res["rpavg"], res["wp"], res["npairs"], res["weightavg"] = r["ravg"], r["cf"], r["npairs"], r["weightavg"]
    if not c_api_timer and not c_cell_timer:
        return res
    return res, (r["api_time"] if c_api_timer else None), None


def xi(boxsize, nthreads, binfile, X, Y, Z, weights=None, weight_type=None, verbose=False, output_ravg=False,
       xbin_refine_factor=2, ybin_refine_factor=2, zbin_refine_factor=1, max_cells_per_dim=100,
       copy_particles=True, enable_min_sep_opt=True, c_api_timer=False, isa="fastest"):
    """3-D correlation function xi(r) in a periodic cube (analytic randoms)."""
    (X, Y, Z), weights, _, dtype = native_inputs((X, Y, Z), weights, None, X, None, weight_type, True)
    opt = _options(dtype, periodic=True, boxsize=boxsize, verbose=verbose, need_avg=output_ravg,
                   refine=(xbin_refine_factor, ybin_refine_factor, zbin_refine_factor), default_refine=(2, 2, 1),
                   max_cells_per_dim=max_cells_per_dim, copy_particles=copy_particles,
                   enable_min_sep_opt=enable_min_sep_opt, c_api_timer=c_api_timer, isa=isa)
    r = _capi.call_xi(_lib.load(), boxsize, nthreads, binfile, X, Y, Z, w=_w0(weights), weight_type=weight_type,
                      options=opt, dtype=dtype)
    res = np.zeros(r["npairs"].size, dtype=[("rmin", np.float64), ("rmax", np.float64), ("ravg", np.float64),
                                            ("xi", np.float64), ("npairs", np.uint64), ("weightavg", np.float64)])
    res["rmin"], res["rmax"] = r["rupp"][:-1], r["rupp"][1:]
    res["ravg"], res["xi"], res["npairs"], res["weightavg"] = r["ravg"], r["cf"], r["npairs"], r["weightavg"]
    return (res, r["api_time"]) if c_api_timer else res


def vpf(rmax, nbins, nspheres, numpN, seed, X, Y, Z, verbose=False, periodic=True, boxsize=None, xbin_refine_factor=1,
        ybin_refine_factor=1, zbin_refine_factor=1, max_cells_per_dim=100, copy_particles=True, c_api_timer=False,
        isa="fastest"):
    """Counts-in-spheres in a simulation box (``Corrfunc.theory.vpf``, Corrfunc/theory/vpf.py:17-200): the probability
    pN that a sphere of radius r holds exactly N points, for ``nbins`` radii up to ``rmax`` and N < ``numpN``, from
    ``nspheres`` centres drawn with MT19937 seeded by ``seed``.  Returns a structured array (rmax, pN[numpN])."""
    if periodic and boxsize is None:
        raise ValueError("Must specify a boxsize if periodic=True")
    X, Y, Z = (convert_to_native_endian(a, warn=True) for a in (X, Y, Z))
    dtype = check_same_dtype(X, Y, Z)
    opt = _options(dtype, periodic=periodic, boxsize=boxsize, verbose=verbose, need_avg=False,
                   refine=(xbin_refine_factor, ybin_refine_factor, zbin_refine_factor), default_refine=(1, 1, 1),
                   max_cells_per_dim=max_cells_per_dim, copy_particles=copy_particles, enable_min_sep_opt=True,
                   c_api_timer=c_api_timer, isa=isa)
    r = _capi.call_vpf(_lib.load(), rmax, nbins, nspheres, numpN, seed, X, Y, Z, options=opt, dtype=dtype)
    res = np.zeros(r["nbin"], dtype=[("rmax", np.float64), ("pN", (np.float64, numpN))])
    res["rmax"] = (np.arange(r["nbin"]) + 1) * (rmax / float(nbins))  # _countpairs.c: r = (ibin + 1) * rstep
    res["pN"] = r["pN"] if numpN > 1 else r["pN"][:, 0]
    return (res, r["api_time"]) if c_api_timer else res
