"""Catalogue readers: fast-food binary and text, against files written here in the layout of
io/io.c:29-283 / docs/source/modules/fast_food_binary.rst, and against the reference's bundled Mr19 mock when
/root/reference is present (its content is also committed as tests/golden/Mr19_mock_northonly_radecw.npz)."""
import os
import struct

import numpy as np
import pytest

import harness as H
from corrfunc_b200.io import read_ascii_catalog, read_catalog, read_fastfood_catalog


def _write_ff(path, cols, width):
    dt = np.float32 if width == 4 else np.float64
    n = cols[0].size
    with open(path, "wb") as f:
        def rec(payload):
            f.write(struct.pack("@i", len(payload)))
            f.write(payload)
            f.write(struct.pack("@i", len(payload)))
        rec(struct.pack("@iiiii", 0, n, 0, 0, 0))
        rec(struct.pack("@fffffffff", *([0.0] * 9)))
        rec(struct.pack("@f", 0.0))
        for c in cols:
            rec(np.asarray(c, dtype=dt).tobytes())


@pytest.mark.parametrize("width", [4, 8])
@pytest.mark.parametrize("ret", [np.float32, np.float64, None])
def test_fastfood_round_trip(tmp_path, width, ret):
    rng = np.random.default_rng(0)
    cols = [rng.random(1000) * 420.0 for _ in range(4)]
    p = str(tmp_path / "cat.ff")
    _write_ff(p, cols, width)
    x, y, z = read_fastfood_catalog(p, ret)
    stored = [np.asarray(c, dtype=np.float32 if width == 4 else np.float64) for c in cols]
    want_dt = np.float64 if ret is None else ret
    for got, s in zip((x, y, z), stored):
        assert got.dtype == want_dt and np.array_equal(got, s.astype(want_dt))
    x, y, z, w = read_fastfood_catalog(p, ret, need_weights=True)
    assert np.array_equal(w, stored[3].astype(want_dt))
    x2, y2, z2 = read_catalog(p, want_dt)
    assert np.array_equal(x2, x) and np.array_equal(z2, z)


def test_ascii_and_errors(tmp_path):
    rng = np.random.default_rng(1)
    a = rng.random((50, 4)) * 100.0
    p = str(tmp_path / "cat.dat")
    np.savetxt(p, a, fmt="%.17g")
    x, y, z = read_ascii_catalog(p)
    assert np.array_equal(x, a[:, 0]) and np.array_equal(y, a[:, 1]) and np.array_equal(z, a[:, 2])
    x32, _, _ = read_catalog(p, np.float32)
    assert x32.dtype == np.float32 and np.array_equal(x32, a[:, 0].astype(np.float32))
    with pytest.raises(IOError):
        read_catalog(str(tmp_path / "missing.ff"))
    with pytest.raises(ValueError):
        read_fastfood_catalog(p, np.int32)
    bad = str(tmp_path / "bad.ff")
    open(bad, "wb").write(struct.pack("@iiiiiii", 16, 0, 5, 0, 0, 0, 16))
    with pytest.raises(AssertionError):
        read_fastfood_catalog(bad)


@pytest.mark.skipif(not os.path.exists("/root/reference/mocks/tests/data/Mr19_mock_northonly.rdcz.ff"),
                    reason="the reference tree is only present in the build container")
def test_reads_the_reference_mock_catalogue():
    ra, dec, cz, w = read_fastfood_catalog("/root/reference/mocks/tests/data/Mr19_mock_northonly.rdcz.ff",
                                           np.float64, need_weights=True)
    gra, gdec, gw = H.load_mr19_mock()
    assert np.array_equal(ra, gra) and np.array_equal(dec, gdec) and np.array_equal(w, gw)
    assert cz.size == ra.size == 84383 and cz.min() > 0
