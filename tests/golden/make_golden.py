"""Generates the committed golden fixtures from /root/reference (run in the build container only;
the GPU box has no /root/reference).  Usage: python tests/golden/make_golden.py

Fixtures:
  Mr19_mock_northonly_radecw.npz  <- mocks/tests/data/Mr19_mock_northonly.rdcz.ff (fast-food binary,
                                     docs/source/modules/fast_food_binary.rst): RA, DEC, cz, weight as float64
  Mr19_mock_wtheta_DD.txt         <- mocks/tests/Mr19_mock_wtheta.DD (npairs thetaavg thetamin thetamax weightavg)
  Mr19_mock_DDrppi_DD.txt         <- mocks/tests/Mr19_mock.DD (DDrppi_mocks autocorr, cz input: npairs rpavg . pi_upper weightavg)
  mocks_bins.txt                  <- mocks/tests/bins
  Mr19_centers_xyz_forVPF_rmax_10Mpc.txt, Mr19_mock_vpf.txt <- mocks/tests/data/..., mocks/tests/Mr19_mock_vpf (vpf_mocks)
  angular_bins.txt                <- mocks/tests/angular_bins
  theory_bins.txt                 <- theory/tests/bins (14 log bins 0.1675-23.8755)
  ref_synthetic_*.npz             <- outputs of the UNMODIFIED reference (oracle/_ref, AVX-512 kernels) on small
                                     seeded synthetic inputs, one per statistic and precision
"""
import os
import shutil
import struct
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
REF = os.environ.get("CORRFUNC_REFERENCE", "/root/reference")


def read_fastfood(filename, nfields=4):
    """Fortran-record fast-food reader (same layout Corrfunc/io.py:60-140 parses)."""
    with open(filename, "rb") as f:
        (s1,) = struct.unpack("@i", f.read(4))
        idat = struct.unpack("@iiiii", f.read(20))
        (s2,) = struct.unpack("@i", f.read(4))
        assert s1 == 20 and s2 == 20
        ngal = idat[1]
        f.seek(4 + 36 + 4 + 4 + 4 + 4, 1)  # fdat + znow records
        out = []
        for _ in range(nfields):
            (s1,) = struct.unpack("@i", f.read(4))
            assert s1 in (ngal * 4, ngal * 8)
            dt = np.float32 if s1 // ngal == 4 else np.float64
            out.append(np.fromfile(f, dt, ngal).astype(np.float64))
            f.read(4)
    return out


def synthetic_inputs(seed, n, boxsize, dtype):
    rng = np.random.default_rng(seed)
    pos = (rng.random((3, n)) * boxsize).astype(dtype)
    w = (1.0 - rng.random(n)).astype(dtype)
    return pos[0], pos[1], pos[2], w


def main():
    ra, dec, cz, w = read_fastfood(os.path.join(REF, "mocks/tests/data/Mr19_mock_northonly.rdcz.ff"))
    np.savez_compressed(os.path.join(HERE, "Mr19_mock_northonly_radecw.npz"), ra=ra, dec=dec, w=w, cz=cz)
    # the reference's own known-answer file for DDrppi_mocks on this catalogue (cz input, cosmology 1, pimax 40;
    # Corrfunc/tests/test_mocks.py:15-34) and its rp bins
    shutil.copy(os.path.join(REF, "mocks/tests/Mr19_mock.DD"), os.path.join(HERE, "Mr19_mock_DDrppi_DD.txt"))
    shutil.copy(os.path.join(REF, "mocks/tests/bins"), os.path.join(HERE, "mocks_bins.txt"))
    # the reference's known-answer test for vpf_mocks (Corrfunc/tests/test_mocks.py:82-110): sphere centres + pN table
    shutil.copy(os.path.join(REF, "mocks/tests/data/Mr19_centers_xyz_forVPF_rmax_10Mpc.txt"), HERE)
    shutil.copy(os.path.join(REF, "mocks/tests/Mr19_mock_vpf"), os.path.join(HERE, "Mr19_mock_vpf.txt"))
    shutil.copy(os.path.join(REF, "mocks/tests/Mr19_mock_wtheta.DD"), os.path.join(HERE, "Mr19_mock_wtheta_DD.txt"))
    shutil.copy(os.path.join(REF, "mocks/tests/angular_bins"), os.path.join(HERE, "angular_bins.txt"))
    shutil.copy(os.path.join(REF, "theory/tests/bins"), os.path.join(HERE, "theory_bins.txt"))

    # reference outputs on seeded synthetic inputs (inputs are regenerated from the seed by the tests)
    import harness as H
    from corrfunc_b200 import _capi as capi

    ref = H.load_ref()
    assert ref is not None, "run `make -C oracle ref` first"
    isa = H.ref_isa()
    bins = np.loadtxt(os.path.join(HERE, "theory_bins.txt"))
    edges = np.concatenate([bins[:1, 0], bins[:, 1]])
    L, N, seed = 420.0, 40000, 2024
    for dtype in (np.float64, np.float32):
        x, y, z, wt = synthetic_inputs(seed, N, L, dtype)
        x2, y2, z2, wt2 = synthetic_inputs(seed + 1, N // 2, L, dtype)
        tag = np.dtype(dtype).name
        out = {}

        def opts(**kw):
            return capi.default_options(dtype, need_avg_sep=True, isa=isa, **kw)

        kw = dict(w1=wt, weight_type="pair_product")
        kwx = dict(w1=wt, weight_type="pair_product", X2=x2, Y2=y2, Z2=z2, w2=wt2)
        for periodic in (True, False):
            p = "per" if periodic else "nonper"
            r = capi.call_DD(ref, 1, 4, edges, x, y, z, options=opts(periodic=periodic, boxsize=L), **kw)
            out["DD_auto_%s" % p] = r
            r = capi.call_DD(ref, 0, 4, edges, x, y, z, options=opts(periodic=periodic, boxsize=L), **kwx)
            out["DD_cross_%s" % p] = r
            r = capi.call_DDrppi(ref, 1, 4, 40.0, edges, x, y, z, options=opts(periodic=periodic, boxsize=L), **kw)
            out["DDrppi_auto_%s" % p] = r
            r = capi.call_DDsmu(ref, 1, 4, edges, 0.5, 10, x, y, z, options=opts(periodic=periodic, boxsize=L), **kw)
            out["DDsmu_auto_%s" % p] = r
        out["xi"] = capi.call_xi(ref, L, 4, edges, x, y, z, w=wt, weight_type="pair_product", options=opts())
        out["wp"] = capi.call_wp(ref, L, 4, 40.0, edges, x, y, z, w=wt, weight_type="pair_product", options=opts())
        flat = {}
        for k, r in out.items():
            for f in ("npairs", "ravg", "weightavg", "cf"):
                if f in r:
                    flat["%s__%s" % (k, f)] = np.asarray(r[f])
        np.savez_compressed(os.path.join(HERE, "ref_synthetic_%s.npz" % tag), seed=seed, N=N, L=L, edges=edges, **flat)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
