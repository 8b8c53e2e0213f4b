"""Test-side loaders for the checkers: the CPU oracle (oracle/libpairs_oracle.so) and the unmodified
reference built into oracle/_ref by oracle/build_ref.sh.  TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")

MODES = dict(DD=0, xi=1, DDrppi=2, wp=3, DDsmu=4, DDtheta=5, DDrppi_mocks=6, DDsmu_mocks=7)


def _cpu_flags():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    return set(line.split(":", 1)[1].split())
    except OSError:
        pass
    return set()


def ref_variant():
    return "v4" if "avx512f" in _cpu_flags() else "v3"


def load_ref():
    """The unmodified reference (AVX-512 kernels when the host has them).  None if not built."""
    path = os.path.join(ORACLE_DIR, "_ref", "libcorrfunc_ref_%s.so" % ref_variant())
    if not os.path.exists(path):
        return None
    return C.CDLL(path, mode=os.RTLD_LOCAL)


def ref_isa():
    # -1 would hit the reference's NULL-kernel static-cache quirk (countpairs_impl.c.src:42-46)
    return 9 if ref_variant() == "v4" else 7


def load_oracle():
    path = os.path.join(ORACLE_DIR, "libpairs_oracle.so")
    if not os.path.exists(path):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "libpairs_oracle.so"])
    lib = C.CDLL(path, mode=os.RTLD_LOCAL)
    return lib


def _p(a):
    return C.c_void_p(a.ctypes.data) if a is not None else C.c_void_p(None)


def oracle_theory(stat, X1, Y1, Z1, bins, *, w1=None, X2=None, Y2=None, Z2=None, w2=None, autocorr=True,
                  pimax=0.0, mu_max=0.0, nmu_bins=0, periodic=True, boxsize=None, refine=(2, 2, 1),
                  custom_refine=False, max_cells=100, enable_min_sep=True, need_avg=False, weight_type=None,
                  nthreads=None):
    lib = load_oracle()
    if nthreads:
        lib.oracle_set_num_threads(int(nthreads))
    dtype = np.asarray(X1).dtype
    fn = lib.oracle_theory_double if dtype == np.float64 else lib.oracle_theory_float
    fn.restype = C.c_int
    arrs = [None if a is None else np.ascontiguousarray(a, dtype=dtype) for a in (X1, Y1, Z1, w1, X2, Y2, Z2, w2)]
    X1, Y1, Z1, w1, X2, Y2, Z2, w2 = arrs
    edges = np.sort(np.asarray(bins, dtype=np.float64))
    nbin = edges.size  # reference's nbin = nlines+1 = number of edges
    mode = MODES[stat]
    if stat in ("DDrppi", "DDrppi_mocks"):
        nslots = (nbin + 1) * (int(pimax) + 1)
    elif stat in ("DDsmu", "DDsmu_mocks"):
        nslots = (nbin + 1) * (nmu_bins + 1)
    else:
        nslots = nbin
    npairs = np.zeros(nslots, dtype=np.uint64)
    avg = np.zeros(nslots)
    wavg = np.zeros(nslots)
    cf = np.zeros(nbin)
    lat = np.zeros(6, dtype=np.int32)
    if boxsize is None:
        bx = (-2.0, -2.0, -2.0)
    else:
        b = np.atleast_1d(np.asarray(boxsize, dtype=np.float64))
        bx = (float(b[0]),) * 3 if b.size == 1 else tuple(float(v) for v in b[:3])
    rf = np.asarray(refine, dtype=np.int32)
    need_w = weight_type is not None
    st = fn(C.c_int(mode), C.c_int64(X1.size), _p(X1), _p(Y1), _p(Z1), _p(w1),
            C.c_int64(0 if X2 is None else X2.size), _p(X2), _p(Y2), _p(Z2), _p(w2), C.c_int(int(autocorr)),
            C.c_int(nbin), _p(edges), C.c_double(pimax), C.c_double(mu_max), C.c_int(nmu_bins),
            C.c_int(int(periodic)), C.c_double(bx[0]), C.c_double(bx[1]), C.c_double(bx[2]), _p(rf),
            C.c_int(int(custom_refine)), C.c_int(max_cells), C.c_int(int(enable_min_sep)), C.c_int(int(need_avg)),
            C.c_int(int(need_w)), _p(npairs), _p(avg), _p(wavg), _p(cf), _p(lat))
    if st != 0:
        raise RuntimeError("oracle failed")
    if stat in ("DDrppi", "DDrppi_mocks"):
        npi = int(pimax)
        g = lambda a: a.reshape(nbin + 1, npi + 1)[1:nbin, :npi].copy()
        return dict(npairs=g(npairs), ravg=g(avg), weightavg=g(wavg), lattice=lat)
    if stat in ("DDsmu", "DDsmu_mocks"):
        g = lambda a: a.reshape(nbin + 1, nmu_bins + 1)[1:nbin, :nmu_bins].copy()
        return dict(npairs=g(npairs), ravg=g(avg), weightavg=g(wavg), lattice=lat)
    return dict(npairs=npairs[1:], ravg=avg[1:], weightavg=wavg[1:], cf=cf[1:], lattice=lat)


FLOAT_WINDOW_CASES = {"DD": 0, "xi": 2}  # statistic -> seed (found by search; see make_golden_float_window.py)


def float_window_case(stat):
    """2 M float32 points in a 60000-wide periodic box, bins up to rmax = 1200: coordinates 50x larger than rmax and
    ~2 particles per cell make the reference's float z-window drop one pair on the edge of the last bin."""
    n, L, rmax = 2000000, 60000.0, 1200.0
    rng = np.random.default_rng(FLOAT_WINDOW_CASES[stat])
    x, y, z = [(rng.random(n) * L).astype(np.float32) for _ in range(3)]
    edges = np.logspace(np.log10(rmax / 50), np.log10(rmax), 8)
    return x, y, z, L, edges


def oracle_vpf_mocks(RA, DEC, D, xc, yc, zc, rmax, nbin, num_pN, nthreads=None, dmax_randoms=0.0):
    """Counts-in-spheres oracle: galaxies as RA, DEC (deg) and comoving distance, centres in the shifted Cartesian
    frame of the reference's centres file.  Returns pN[nbin, num_pN] and the shift the reference applies."""
    lib = load_oracle()
    if nthreads:
        lib.oracle_set_num_threads(int(nthreads))
    dtype = np.asarray(RA).dtype
    fn = lib.oracle_vpf_mocks_double if dtype == np.float64 else lib.oracle_vpf_mocks_float
    fn.restype = C.c_int
    RA, DEC, D, xc, yc, zc = [np.ascontiguousarray(a, dtype=dtype) for a in (RA, DEC, D, xc, yc, zc)]
    pN = np.zeros((nbin, num_pN))
    rcube = C.c_double(0.0)
    st = fn(C.c_int64(RA.size), _p(RA), _p(DEC), _p(D), C.c_int64(xc.size), _p(xc), _p(yc), _p(zc), C.c_double(rmax),
            C.c_int(nbin), C.c_int(num_pN), _p(pN), C.c_double(float(dmax_randoms)), C.byref(rcube))
    if st != 0:
        raise RuntimeError("oracle_vpf_mocks failed")
    return pN, rcube.value


def mt19937_uniform(seed, n):
    """The library's restatement of gsl_rng_mt19937 + gsl_rng_uniform (host code, no GPU)."""
    from corrfunc_b200 import _lib

    lib = _lib.load()
    out = np.zeros(n)
    lib.corrfunc_b200_mt19937_uniform.restype = None
    lib.corrfunc_b200_mt19937_uniform(C.c_ulong(seed), C.c_int64(n), out.ctypes.data_as(C.c_void_p))
    return out


def vpf_theory_centres(X, Y, Z, rmax, nc, seed, periodic, boxsize):
    """The sphere centres theory/vpf draws (countspheres_impl.c.src:299-316) from the library's MT19937 stream."""
    dt = X.dtype.type
    lo = [dt(a.min()) for a in (X, Y, Z)]
    hi = [dt(a.max()) for a in (X, Y, Z)]
    wrap = [dt(boxsize) if (periodic and boxsize > 0) else dt(h - l) for l, h in zip(lo, hi)]
    u = mt19937_uniform(seed, 3 * (nc * 50 + 1000))
    cen, t = [], 0
    while len(cen) < nc:
        c = [dt(np.float64(wrap[a]) * u[3 * t + a] + np.float64(lo[a])) for a in range(3)]
        t += 1
        if not periodic and any(np.float64(c[a] - lo[a]) < rmax or np.float64(hi[a] - c[a]) < rmax for a in range(3)):
            continue
        cen.append(c)
    cen = np.array(cen, dtype=X.dtype)
    return cen[:, 0].copy(), cen[:, 1].copy(), cen[:, 2].copy(), [float(w) for w in wrap]


def oracle_vpf_theory(X, Y, Z, xc, yc, zc, periodic, wrap, rmax, nbin, num_pN, nthreads=None):
    lib = load_oracle()
    if nthreads:
        lib.oracle_set_num_threads(int(nthreads))
    dtype = np.asarray(X).dtype
    fn = lib.oracle_vpf_theory_double if dtype == np.float64 else lib.oracle_vpf_theory_float
    fn.restype = C.c_int
    X, Y, Z, xc, yc, zc = [np.ascontiguousarray(a, dtype=dtype) for a in (X, Y, Z, xc, yc, zc)]
    pN = np.zeros((nbin, num_pN))
    st = fn(C.c_int64(X.size), _p(X), _p(Y), _p(Z), C.c_int64(xc.size), _p(xc), _p(yc), _p(zc), C.c_int(int(periodic)),
            C.c_double(wrap[0]), C.c_double(wrap[1]), C.c_double(wrap[2]), C.c_double(rmax), C.c_int(nbin),
            C.c_int(num_pN), _p(pN))
    if st != 0:
        raise RuntimeError("oracle_vpf_theory failed")
    return pN


def mock_points(seed, n, dtype):
    """Seeded synthetic survey wedge: RA 40-90 deg, DEC -10..30 deg, comoving distance 300-700 (uniform in volume
    along the radius), weights in [0.5, 1.5)."""
    rng = np.random.default_rng(seed)
    ra = (40.0 + 50.0 * rng.random(n)).astype(dtype)
    dec = (-10.0 + 40.0 * rng.random(n)).astype(dtype)
    d = (300.0 + 400.0 * rng.random(n) ** (1.0 / 3.0)).astype(dtype)
    w = (0.5 + rng.random(n)).astype(dtype)
    return ra, dec, d, w


def oracle_config(name, literal=0, nthreads=None):
    """A BASELINE config (bench.CONFIGS, bench.py's own seeded input) through the oracle; literal=1 follows the
    AVX-512 kernels' z-sorted chunked control flow for wp / DDrppi (oracle_impl.h)."""
    import bench

    cfg = bench.CONFIGS[name]
    dtype = np.float32 if cfg["dtype"] == "f32" else np.float64
    pts = bench.gen_points(cfg, cfg["N"], dtype)
    lib = load_oracle()
    lib.oracle_set_literal_kernels(int(literal))
    try:
        return oracle_theory(cfg["stat"], pts["x"], pts["y"], pts["z"], bench.make_bins(cfg["bins"]),
                             pimax=cfg.get("pimax", 0.0), periodic=True, boxsize=cfg["L"],
                             nthreads=nthreads or os.cpu_count())
    finally:
        lib.oracle_set_literal_kernels(0)


def oracle_theta(RA1, DEC1, bins, *, w1=None, RA2=None, DEC2=None, w2=None, autocorr=True, link_in_dec=True,
                 link_in_ra=True, ra_refine=2, dec_refine=2, max_cells=100, enable_min_sep=True, need_avg=False,
                 weight_type=None, fast_acos=False, nthreads=None):
    """DDtheta oracle.  RA in [0,360], DEC in [-90,90] expected (the wrappers' fix_ra_dec does that)."""
    lib = load_oracle()
    if nthreads:
        lib.oracle_set_num_threads(int(nthreads))
    dtype = np.asarray(RA1).dtype
    fn = lib.oracle_theta_double if dtype == np.float64 else lib.oracle_theta_float
    fn.restype = C.c_int
    RA1, DEC1, w1, RA2, DEC2, w2 = [None if a is None else np.ascontiguousarray(a, dtype=dtype)
                                    for a in (RA1, DEC1, w1, RA2, DEC2, w2)]
    edges = np.sort(np.asarray(bins, dtype=np.float64))
    if dtype == np.float32:  # setup_bins_float parses the file with %f
        edges = edges.astype(np.float32).astype(np.float64)
    nbin = edges.size
    npairs = np.zeros(nbin, dtype=np.uint64)
    avg = np.zeros(nbin)
    wavg = np.zeros(nbin)
    lat = np.zeros(3, dtype=np.int32)
    need_w = weight_type is not None
    st = fn(C.c_int64(RA1.size), _p(RA1), _p(DEC1), _p(w1), C.c_int64(0 if RA2 is None else RA2.size), _p(RA2),
            _p(DEC2), _p(w2), C.c_int(int(autocorr)), C.c_int(nbin), _p(edges), C.c_int(int(link_in_dec)),
            C.c_int(int(link_in_ra)), C.c_int(ra_refine), C.c_int(dec_refine), C.c_int(max_cells),
            C.c_int(int(enable_min_sep)), C.c_int(int(need_avg)), C.c_int(int(need_w)), C.c_int(int(fast_acos)),
            _p(npairs), _p(avg), _p(wavg), _p(lat))
    if st != 0:
        raise RuntimeError("oracle_theta failed")
    return dict(npairs=npairs[1:], ravg=avg[1:], weightavg=wavg[1:], lattice=lat)


GOLDEN = os.path.join(ROOT, "tests", "golden")


def load_bins_file(name):
    b = np.loadtxt(os.path.join(GOLDEN, name))
    return np.concatenate([b[:1, 0], b[:, 1]])


def load_mr19_mock():
    d = np.load(os.path.join(GOLDEN, "Mr19_mock_northonly_radecw.npz"))
    return d["ra"], d["dec"], d["w"]


def load_mr19_mock_cz():
    d = np.load(os.path.join(GOLDEN, "Mr19_mock_northonly_radecw.npz"))
    return d["ra"], d["dec"], d["cz"], d["w"]


def load_ddrppi_mocks_golden():
    """mocks/tests/Mr19_mock.DD: rows rp-major x 40 pi bins; columns npairs rpavg . pi_upper weightavg."""
    g = np.loadtxt(os.path.join(GOLDEN, "Mr19_mock_DDrppi_DD.txt"))
    return dict(npairs=g[:, 0].astype(np.uint64), ravg=g[:, 1], weightavg=g[:, 4])


VPF_CENTERS = os.path.join(GOLDEN, "Mr19_centers_xyz_forVPF_rmax_10Mpc.txt")


def load_vpf_golden():
    """mocks/tests/Mr19_mock_vpf: one row per radius 1..10: r, p0..p5."""
    return np.genfromtxt(os.path.join(GOLDEN, "Mr19_mock_vpf.txt"), usecols=range(1, 7), dtype=np.float64)


def cz_to_comoving(cz, cosmology):
    """The library's host-side redshift -> distance conversion (no GPU needed)."""
    from corrfunc_b200 import _lib

    lib = _lib.load()
    lib.corrfunc_b200_cz_to_comoving.restype = C.c_int
    out = np.zeros_like(cz)
    st = lib.corrfunc_b200_cz_to_comoving(C.c_int(cz.itemsize), C.c_int64(cz.size), cz.ctypes.data_as(C.c_void_p),
                                          C.c_int(cosmology), out.ctypes.data_as(C.c_void_p))
    if st != 0:
        raise RuntimeError("corrfunc_b200_cz_to_comoving failed")
    return out


def vpf_centres_from_randoms(RA, DEC, D, rcube, rmax, threshold, nc):
    """The reference's choice of sphere centres when it has no usable centres file (countspheres_mocks_impl.c.src:
    140-204, 478-490): the first nc randoms, in input order, with more than `threshold` randoms (itself included)
    within rmax; r2 = dx*dx + dy*dy + dz*dz in the run precision, without FMA.  rcube = the shift (max distance + 1).
    Brute force: small inputs only."""
    dt = RA.dtype.type
    k = 0.017453292519943295769236907684886127134428718885417254560971  # PI_OVER_180 (utils/function_precision.h)

    def cosd(a):  # COSD(x): cos / cosf of the DOUBLE product x * PI_OVER_180, rounded to the run precision first
        return np.cos((a.astype(np.float64) * k).astype(dt))

    def sind(a):
        return np.sin((a.astype(np.float64) * k).astype(dt))

    x = (D * cosd(DEC) * cosd(RA)).astype(dt) + dt(rcube)
    y = (D * cosd(DEC) * sind(RA)).astype(dt) + dt(rcube)
    z = (D * sind(DEC)).astype(dt) + dt(rcube)
    r2max = dt(rmax) * dt(rmax)
    keep = []
    for i in range(x.size):
        dx, dy, dz = x - x[i], y - y[i], z - z[i]
        r2 = dx * dx + dy * dy + dz * dz
        if np.count_nonzero(r2 < r2max) > threshold:
            keep.append(i)
            if len(keep) == nc:
                break
    keep = np.asarray(keep, dtype=np.int64)
    return x[keep], y[keep], z[keep]


def load_wtheta_golden():
    g = np.loadtxt(os.path.join(GOLDEN, "Mr19_mock_wtheta_DD.txt"))
    return dict(npairs=g[:, 0].astype(np.uint64), ravg=g[:, 1], weightavg=g[:, 4])


def sphere_points(seed, n, dtype):
    rng = np.random.default_rng(seed)
    ra = (360.0 * rng.random(n)).astype(dtype)
    dec = np.degrees(np.arcsin(2.0 * rng.random(n) - 1.0)).astype(dtype)
    return ra, dec


def box_points(seed, n, boxsize, dtype):
    rng = np.random.default_rng(seed)
    pos = (rng.random((3, n)) * boxsize).astype(dtype)
    w = (1.0 - rng.random(n)).astype(dtype)
    return pos[0], pos[1], pos[2], w
