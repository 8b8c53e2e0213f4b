"""Drop-in check of the Python API: every wrapper takes the reference's parameters, in the reference's order, with
the reference's defaults (tests/golden/reference_signatures.json was extracted from /root/reference/Corrfunc by
tests/golden/make_golden_signatures.py)."""
import inspect
import json
import os

import pytest

GOLD = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_signatures.json")))


def _ours(name):
    if name.endswith("_mocks"):
        import corrfunc_b200.mocks as M

        return getattr(M, name)
    import corrfunc_b200.utils as U

    if hasattr(U, name):
        return getattr(U, name)
    import corrfunc_b200.theory as T

    return getattr(T, name)


@pytest.mark.parametrize("name", sorted(GOLD))
def test_signature_matches_reference(name):
    sig = inspect.signature(_ours(name))
    params = list(sig.parameters.values())
    want = GOLD[name]
    assert [p.name for p in params] == want["args"], want["source"]
    for p in params:
        if p.name in want["defaults"]:
            assert p.default == want["defaults"][p.name], (p.name, want["source"])
        else:
            assert p.default is inspect.Parameter.empty, (p.name, want["source"])


def test_compat_package_exposes_the_reference_import_paths():
    """`compat/` holds a package named Corrfunc that forwards the reference's import paths to this library."""
    import importlib
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "compat"))
    try:
        for mod, attr in (("Corrfunc.theory.DD", "DD"), ("Corrfunc.theory.DDrppi", "DDrppi"),
                          ("Corrfunc.theory.DDsmu", "DDsmu"), ("Corrfunc.theory.wp", "wp"), ("Corrfunc.theory.xi", "xi"),
                          ("Corrfunc.mocks.DDtheta_mocks", "DDtheta_mocks"), ("Corrfunc.theory", "DD"),
                          ("Corrfunc.mocks", "DDtheta_mocks"), ("Corrfunc.mocks.DDrppi_mocks", "DDrppi_mocks"),
                          ("Corrfunc.mocks.DDsmu_mocks", "DDsmu_mocks"), ("Corrfunc.mocks", "DDsmu_mocks"), ("Corrfunc.mocks.vpf_mocks", "vpf_mocks"), ("Corrfunc.theory.vpf", "vpf"), ("Corrfunc.theory", "vpf"),
                          ("Corrfunc.utils", "convert_3d_counts_to_cf"),
                          ("Corrfunc.utils", "convert_rp_pi_counts_to_wp")):
            m = importlib.import_module(mod)
            assert callable(getattr(m, attr)), (mod, attr)
        import Corrfunc

        assert Corrfunc.__version__
        assert Corrfunc.which("sh") and callable(Corrfunc.read_text_file) and callable(Corrfunc.write_text_file)
        import Corrfunc.io

        assert callable(Corrfunc.io.read_catalog)
    finally:
        sys.path.pop(0)
        for k in [k for k in sys.modules if k == "Corrfunc" or k.startswith("Corrfunc.")]:
            del sys.modules[k]


def test_wrapper_input_preparation_follows_the_reference_order():
    """Corrfunc/theory/DD.py:222-247: weights are shaped first (a Python scalar takes the particles' dtype), then every
    array is brought to native byte order, and only then must all arrays share one dtype."""
    import numpy as np
    from corrfunc_b200.utils import native_inputs

    x = np.linspace(0, 1, 5, dtype=np.float32)
    (x1, y1, z1), w1, w2, dt = native_inputs((x, x, x), 0.5, None, x, None, "pair_product", True)
    assert dt == np.float32 and w1.dtype == np.float32 and w1.shape == (1, 5) and w2 is None
    big = x.astype(">f4")
    (x1, y1, z1), w1, _, dt = native_inputs((big, big, big), big, None, big, None, "pair_product", True)
    assert dt == np.float32 and x1.dtype.isnative and w1.dtype.isnative and np.array_equal(x1, x)
    # cross-correlation with one weight array missing: ones for pair_product
    (a, b, c, d, e, f), w1, w2, dt = native_inputs((x, x, x, x[:3], x[:3], x[:3]), None, x[:3], x, x[:3], "pair_product", False)
    assert w1.shape == (1, 5) and np.all(w1 == 1) and w2.shape == (1, 3)
    import pytest
    with pytest.raises(TypeError):
        native_inputs((x, x.astype(np.float64), x), None, None, x, None, None, True)
