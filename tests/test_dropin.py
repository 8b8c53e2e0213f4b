"""Drop-in proof with the reference's own headers: tests/c_abi/caller.c is an ordinary user program of the reference's
static-library interface (docs/source/staticlibrary-interface.rst:33-117).  tests/c_abi/build.sh compiles it twice --
`caller_ref` against /root/reference's utils/defs.h + theory/*/countpairs*.h + mocks/DDtheta_mocks/countpairs_theta_mocks.h
(build container only; the binary travels to the GPU box), `caller_inc` against this repo's include/ -- and links both
against libcorrfunc_b200.so.  What they print must equal the CPU oracle: npairs bit for bit, averages within tolerance.

Without a GPU the same binaries must fail LOUDLY (no CPU fallback)."""
import os
import subprocess

import numpy as np
import pytest

import harness as H

BUILD = os.path.join(H.ROOT, "tests", "c_abi", "_build")
TOL = {np.float64: 1e-10, np.float32: 1e-5}


def _ensure_built():
    if not os.path.exists(os.path.join(BUILD, "caller_inc")):
        subprocess.check_call(["bash", os.path.join(H.ROOT, "tests", "c_abi", "build.sh")])


def _binaries():
    _ensure_built()
    out = [os.path.join(BUILD, "caller_inc")]
    if os.path.exists(os.path.join(BUILD, "caller_ref")):  # built where /root/reference exists
        out.append(os.path.join(BUILD, "caller_ref"))
    return out


def _write_inputs(tmp_path, dtype, set1, set2, bins):
    pfile, bfile = str(tmp_path / "particles.bin"), str(tmp_path / "bins.txt")
    n1 = set1[0].size
    n2 = 0 if set2 is None else set2[0].size
    with open(pfile, "wb") as f:
        np.array([n1, n2], dtype=np.int64).tofile(f)
        for s in (set1, set2):
            if s is None:
                continue
            for a in s:
                np.ascontiguousarray(a, dtype=dtype).tofile(f)
    with open(bfile, "w") as f:
        for lo, hi in zip(bins[:-1], bins[1:]):
            f.write("%s %s\n" % (repr(float(lo)), repr(float(hi))))
    return pfile, bfile


def _run(binary, stat, dtype, pfile, bfile, boxsize, *extra, env=None):
    cmd = [binary, stat, str(np.dtype(dtype).itemsize), pfile, bfile, repr(float(boxsize))] + [str(e) for e in extra]
    p = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=600)
    return p.returncode, np.array([[float(v) for v in line.split()] for line in p.stdout.splitlines() if line.strip()]), p.stderr


def _close(a, b, tol, what):
    a, b = np.asarray(a, dtype=np.float64).ravel(), np.asarray(b, dtype=np.float64).ravel()
    rel = np.abs(a - b) / np.maximum(np.abs(b), 1e-300)
    rel[(a == 0) & (b == 0)] = 0
    assert rel.max() <= tol, "%s: max rel diff %.3e" % (what, rel.max())


def test_callers_fail_loudly_without_a_gpu(tmp_path):
    """CPU box: the binaries link, load and run -- and every entry point refuses to count without a CUDA device."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    x, y, z, w = H.box_points(1, 1000, 50.0, np.float64)
    bins = np.linspace(0.5, 5.0, 6)
    pfile, bfile = _write_inputs(tmp_path, np.float64, (x, y, z, w), None, bins)
    for binary in _binaries():
        rc, out, err = _run(binary, "xi", np.float64, pfile, bfile, 50.0)
        assert rc == 1 and out.size == 0
        assert "no CUDA device" in err or "CUDA" in err


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("stat", ["DD", "DDx", "DDrppi", "DDsmu", "wp", "xi", "DDtheta"])
def test_reference_header_caller_matches_the_oracle(tmp_path, stat, dtype):
    L, N = 120.0, 40000
    x, y, z, w = H.box_points(91, N, L, dtype)
    bins = np.logspace(-0.5, np.log10(12.0), 11)
    set2 = None
    okw = dict(w1=w, weight_type="pair_product", need_avg=True, periodic=True, boxsize=L)
    extra = ()
    name = stat
    if stat == "DDx":  # cross-correlation
        name = "DD"
        x2, y2, z2, w2 = H.box_points(92, 25000, L, dtype)
        set2 = (x2, y2, z2, w2)
        okw.update(X2=x2, Y2=y2, Z2=z2, w2=w2, autocorr=False)
    if stat in ("DDrppi", "wp"):
        extra = (30.0,)
        okw.update(pimax=30.0)
    if stat == "DDsmu":
        extra = (0.8, 10)
        okw.update(mu_max=0.8, nmu_bins=10)
    if stat == "DDtheta":
        ra, dec = H.sphere_points(93, 30000, dtype)
        wt = (1.0 - np.random.default_rng(94).random(ra.size)).astype(dtype)
        bins = np.logspace(np.log10(0.05), 1, 13)
        ref = H.oracle_theta(ra, dec, bins, w1=wt, weight_type="pair_product", need_avg=True)
        pfile, bfile = _write_inputs(tmp_path, dtype, (ra, dec, np.zeros_like(ra), wt), None, bins)
        boxsize = 0.0
    else:
        ref = H.oracle_theory(name, x, y, z, bins, **okw)
        pfile, bfile = _write_inputs(tmp_path, dtype, (x, y, z, w), set2, bins)
        boxsize = L
    for binary in _binaries():
        rc, out, err = _run(binary, name, dtype, pfile, bfile, boxsize, *extra)
        assert rc == 0, err
        assert np.array_equal(out[:, 0].astype(np.uint64), ref["npairs"].ravel().astype(np.uint64)), os.path.basename(binary)
        _close(out[:, 1], ref["ravg"], 1e-9 if (stat == "DDtheta" and dtype == np.float64) else (1e-4 if stat == "DDtheta" else TOL[dtype]),
               "average separation (%s)" % os.path.basename(binary))
        _close(out[:, 2], ref["weightavg"], TOL[dtype], "weightavg")
        if stat in ("xi", "wp"):
            _close(out[:, 3], ref["cf"], 1e-9 if dtype == np.float64 else 1e-4, stat)
        assert "c_api_time" in err


@pytest.mark.gpu
def test_one_call_many_gpus(tmp_path):
    """CORRFUNC_B200_NGPUS: one countpairs_xi() call of the C program sharded over every visible device inside the
    library (host thread per device, NVLink peer replication, histograms summed on the host).  Counts must not depend on
    the number of devices.  Runs with whatever the box has (1 device: the single-device path twice)."""
    import torch

    ndev = torch.cuda.device_count()
    L, N = 300.0, 600000
    x, y, z, w = H.box_points(95, N, L, np.float32)
    bins = np.logspace(-0.5, np.log10(20.0), 13)
    pfile, bfile = _write_inputs(tmp_path, np.float32, (x, y, z, w), None, bins)
    binary = _binaries()[-1]
    env1 = dict(os.environ, CORRFUNC_B200_NGPUS="1")
    envn = dict(os.environ, CORRFUNC_B200_NGPUS=str(max(ndev, 1)))
    envn.pop("CORRFUNC_B200_DEVICE", None)
    rc1, out1, err1 = _run(binary, "xi", np.float32, pfile, bfile, L, env=env1)
    rcn, outn, errn = _run(binary, "xi", np.float32, pfile, bfile, L, env=envn)
    assert rc1 == 0 and rcn == 0, err1 + errn
    assert np.array_equal(out1[:, 0], outn[:, 0])
    _close(outn[:, 1], out1[:, 1], 1e-6, "ravg")
    ref = H.oracle_theory("xi", x, y, z, bins, boxsize=L)
    assert np.array_equal(out1[:, 0].astype(np.uint64), ref["npairs"])


def _load_ext(name):
    import importlib.util

    path = os.path.join(BUILD, name + ".so")
    if not os.path.exists(path):
        pytest.skip("%s.so is built only where /root/reference exists (tests/c_abi/build.sh)" % name)
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_reference_extension_modules_import():
    """The reference's CPython extension modules (theory/python_bindings/_countpairs.c,
    mocks/python_bindings/_countpairs_mocks.c), compiled UNMODIFIED from where they lie and linked against
    libcorrfunc_b200.so: every symbol they need resolves (incl. get_max_usable_isa, countspheres)."""
    _ensure_built()
    th = _load_ext("_countpairs")
    mk = _load_ext("_countpairs_mocks")
    assert {"countpairs", "countpairs_rp_pi", "countpairs_s_mu", "countpairs_wp", "countpairs_xi", "countspheres_vpf"} <= set(dir(th))
    assert {"countpairs_rp_pi_mocks", "countpairs_s_mu_mocks", "countpairs_theta_mocks", "countspheres_vpf_mocks"} <= set(dir(mk))


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_reference_extension_modules_count_on_the_gpu(tmp_path, dtype):
    """Calls them the way Corrfunc/theory/DD.py:249-264, DDrppi.py, DDsmu.py, wp.py, xi.py and
    Corrfunc/mocks/DDtheta_mocks.py do (same positional and keyword arguments, weights as a (1, N) array, boxsize as a
    3-tuple, isa = -1 "fastest") and compares the rows they return with the oracle."""
    th = _load_ext("_countpairs")
    mk = _load_ext("_countpairs_mocks")
    L, N = 120.0, 40000
    x, y, z, w = H.box_points(191, N, L, dtype)
    bins = np.logspace(-0.5, np.log10(12.0), 11)
    _, bfile = _write_inputs(tmp_path, dtype, (x, y, z, w), None, bins)
    W = w.reshape(1, -1)
    kw = dict(periodic=True, verbose=False, boxsize=(L, L, L), xbin_refine_factor=2, ybin_refine_factor=2,
              zbin_refine_factor=1, max_cells_per_dim=100, copy_particles=True, enable_min_sep_opt=True, c_api_timer=True,
              isa=-1, weights1=W, weight_type="pair_product")
    okw = dict(w1=w, weight_type="pair_product", need_avg=True, periodic=True, boxsize=L)
    tol = TOL[dtype]

    def rows(res):
        assert res is not None, "the extension returned None (RuntimeError in the reference's wrapper)"
        r, api_time = res
        assert api_time > 0
        return np.array([tuple(t) for t in r], dtype=np.float64)

    # DD: rows (rmin, rmax, ravg, npairs, weightavg), _countpairs.c:1426-1440
    r = rows(th.countpairs(1, 2, bfile, x, y, z, output_ravg=True, **kw))
    ref = H.oracle_theory("DD", x, y, z, bins, **okw)
    assert np.array_equal(r[:, 3].astype(np.uint64), ref["npairs"])
    _close(r[:, 2], ref["ravg"], tol, "ravg")
    _close(r[:, 4], ref["weightavg"], tol, "weightavg")
    # DDrppi: rows (rmin, rmax, rpavg, pi_upper, npairs, weightavg), :1722-1740
    r = rows(th.countpairs_rp_pi(1, 2, 30.0, bfile, x, y, z, output_rpavg=True, **kw))
    ref = H.oracle_theory("DDrppi", x, y, z, bins, pimax=30.0, **okw)
    assert np.array_equal(r[:, 4].astype(np.uint64), ref["npairs"].ravel())
    _close(r[:, 5], ref["weightavg"].ravel(), tol, "weightavg")
    # DDsmu: rows (smin, smax, savg, mu_upper, npairs, weightavg), :2508
    r = rows(th.countpairs_s_mu(1, 2, bfile, 0.8, 10, x, y, z, output_savg=True, fast_divide_and_NR_steps=0, **kw))
    ref = H.oracle_theory("DDsmu", x, y, z, bins, mu_max=0.8, nmu_bins=10, **okw)
    assert np.array_equal(r[:, 4].astype(np.uint64), ref["npairs"].ravel())
    _close(r[:, 2], ref["ravg"].ravel(), tol, "savg")
    # xi: rows (rmin, rmax, ravg, xi, npairs, weightavg), :2201
    kx = {k: v for k, v in kw.items() if k not in ("periodic", "boxsize", "weights1")}
    r = rows(th.countpairs_xi(L, 2, bfile, x, y, z, weights=W, output_ravg=True, **kx))
    ref = H.oracle_theory("xi", x, y, z, bins, **okw)
    assert np.array_equal(r[:, 4].astype(np.uint64), ref["npairs"])
    assert np.allclose(r[:, 3], ref["cf"], rtol=1e-9 if dtype == np.float64 else 1e-4, atol=1e-12)
    # wp: rows (rmin, rmax, rpavg, wp, npairs, weightavg) + cell timings, :1965-1983
    res = th.countpairs_wp(L, 30.0, 2, bfile, x, y, z, weights=W, output_rpavg=True, c_cell_timer=False, **kx)
    assert res is not None
    r = np.array([tuple(t) for t in res[0]], dtype=np.float64)
    ref = H.oracle_theory("wp", x, y, z, bins, pimax=30.0, **okw)
    assert np.array_equal(r[:, 4].astype(np.uint64), ref["npairs"])
    # DDtheta_mocks: rows (thetamin, thetamax, thetaavg, npairs, weightavg), _countpairs_mocks.c:2018
    ra, dec = H.sphere_points(193, 30000, dtype)
    wt = (1.0 - np.random.default_rng(194).random(ra.size)).astype(dtype)
    tb = np.logspace(np.log10(0.05), 1, 13)
    _, tfile = _write_inputs(tmp_path, dtype, (ra, dec, ra, wt), None, tb)
    r = rows(mk.countpairs_theta_mocks(1, 2, tfile, ra, dec, weights1=wt.reshape(1, -1), weight_type="pair_product",
                                       link_in_dec=True, link_in_ra=True, verbose=False, output_thetaavg=True, fast_acos=False,
                                       ra_refine_factor=2, dec_refine_factor=2, max_cells_per_dim=100, copy_particles=True,
                                       enable_min_sep_opt=True, c_api_timer=True, isa=-1))
    ref = H.oracle_theta(ra, dec, tb, w1=wt, weight_type="pair_product", need_avg=True)
    assert np.array_equal(r[:, 3].astype(np.uint64), ref["npairs"])
    _close(r[:, 4], ref["weightavg"], tol, "weightavg")
