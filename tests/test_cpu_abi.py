"""CPU tests of the drop-in boundary: struct layouts, exported symbols, loud failure without a GPU."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import harness as H
from corrfunc_b200 import _capi, _lib

INCLUDE = os.path.join(H.ROOT, "include")


def test_struct_sizes_match_reference_abi():
    assert C.sizeof(_capi.ConfigOptions) == 1024  # utils/defs.h:50 OPTIONS_HEADER_SIZE
    assert C.sizeof(_capi.ExtraOptions) == 1024   # utils/defs.h:349
    assert _capi.ConfigOptions.float_type.offset == 104  # 11 doubles + pointer + int64
    assert _capi.ConfigOptions.version.offset == 116


def test_headers_compile_as_c_and_cxx_and_offsets_agree(tmp_path):
    src = tmp_path / "t.c"
    src.write_text('#include <stddef.h>\n#include <stdio.h>\n#include "countpairs.h"\n#include "countpairs_rp_pi.h"\n'
                   '#include "countpairs_s_mu.h"\n#include "countpairs_wp.h"\n#include "countpairs_xi.h"\n'
                   '#include "countpairs_theta_mocks.h"\n#include "countpairs_rp_pi_mocks.h"\n#include "countpairs_s_mu_mocks.h"\n'
                   '#include "countspheres_mocks.h"\n#include "countspheres.h"\n'
                   '#include "corrfunc_b200.h"\n'
                   'int main(void){struct config_options o=get_config_options();'
                   'printf("%zu %zu %zu %zu %zu %zu %s\\n",sizeof(struct config_options),sizeof(struct extra_options),'
                   'offsetof(struct config_options,float_type),offsetof(struct config_options,version),'
                   'offsetof(struct config_options,bin_refine_factors),offsetof(struct config_options,binning_flags),o.version);return 0;}\n')
    exe = tmp_path / "t"
    subprocess.check_call(["/usr/bin/gcc", "-std=c11", "-I", INCLUDE, str(src), "-o", str(exe)])
    out = subprocess.check_output([str(exe)], text=True).split()
    assert out[:2] == ["1024", "1024"]
    assert int(out[2]) == _capi.ConfigOptions.float_type.offset
    assert int(out[3]) == _capi.ConfigOptions.version.offset
    assert int(out[4]) == _capi.ConfigOptions.bin_refine_factors.offset
    assert int(out[5]) == _capi.ConfigOptions.binning_flags.offset
    assert out[6] == "2.5.3"
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-fsyntax-only", "-x", "c++", "-I", INCLUDE,
                           os.path.join(INCLUDE, "countpairs.h")])


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    declared = set()
    for fn in os.listdir(INCLUDE):
        txt = open(os.path.join(INCLUDE, fn)).read()
        declared |= set(re.findall(r"\b(countpairs\w*|countspheres\w*|free_results\w*|corrfunc_b200_\w+|cfb_\w+)\s*\(", txt))
    declared -= {"cfb_mode"}
    assert set(_capi.EXPORTED_SYMBOLS) <= declared
    for sym in sorted(declared):
        assert hasattr(lib, sym), "libcorrfunc_b200.so does not export %s" % sym


def test_product_never_touches_the_oracle():
    """The product tree must not import, link or dlopen anything under oracle/."""
    pkg = os.path.join(H.ROOT, "corrfunc_b200")
    banned = ("pairs_oracle", "libcorrfunc_ref", "oracle_theory", "oracle_theta", "import harness", "oracle/", "_ref/")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".c", ".h", ".cu", ".cuh", "Makefile")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                for b in banned:
                    assert b not in txt, "%s mentions %s" % (f, b)
    out = subprocess.check_output(["ldd", _lib.LIB_PATH], text=True)
    assert "oracle" not in out and "corrfunc_ref" not in out


def test_no_gpu_means_loud_failure():
    try:
        import torch

        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    import corrfunc_b200.theory as T

    x = np.random.default_rng(0).random(100) * 10
    with pytest.raises(RuntimeError):
        T.DD(1, 1, np.linspace(0.1, 2, 5), x, x, x, boxsize=10.0)


def test_bad_inputs_rejected_before_the_device():
    import corrfunc_b200.theory as T

    x = np.zeros(10)
    with pytest.raises(ValueError):
        T.DD(0, 1, [0.1, 1.0], x, x, x, boxsize=10.0)  # cross without second set
    with pytest.raises(ValueError):
        T.DD(1, 1, [0.1, 1.0], x, x, x)  # periodic without boxsize
    with pytest.raises(TypeError):
        T.DD(1, 1, [0.1, 1.0], x, x.astype(np.float32), x, boxsize=10.0)
    with pytest.raises(ValueError):
        T.DDsmu(1, 1, [0.1, 1.0], 1.5, 10, x, x, x, boxsize=10.0)
    import corrfunc_b200.mocks as M

    with pytest.raises(ValueError):
        M.DDsmu_mocks(0, 1, 1, 0.5, 4, [0.1, 1.0], x, x, x + 100.0, is_comoving_dist=True)  # cross without second set
    with pytest.raises(ValueError):
        M.DDsmu_mocks(1, 1, 1, 1.5, 4, [0.1, 1.0], x, x, x + 100.0, is_comoving_dist=True)  # mu_max > 1
    o = _capi.default_options(np.float64, boxsize=10.0)
    o.version = b"1.0.0"
    with pytest.raises(RuntimeError):
        _capi.call_DD(_lib.load(), 1, 1, [0.1, 1.0], x, x, x, options=o)
