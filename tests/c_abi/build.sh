#!/bin/bash
# Builds tests/c_abi/_build/caller_ref: caller.c compiled against the REFERENCE's own headers (not include/) and linked
# against libcorrfunc_b200.so -- the drop-in proof of INTEGRATION.md section 2.  Build container only (needs
# /root/reference); the binary travels to the GPU box with the snapshot.  A twin compiled against include/ (caller_inc)
# is built everywhere.
set -e
cd "$(dirname "$0")"
ROOT=$(cd ../.. && pwd)
REF=${REF:-/root/reference}
mkdir -p _build
LINK="-L$ROOT/corrfunc_b200/csrc -lcorrfunc_b200 -Wl,-rpath,$ROOT/corrfunc_b200/csrc -Wl,-rpath,\$ORIGIN/../../../corrfunc_b200/csrc -lm"
if [ -d "$REF/utils" ]; then
  gcc -std=c99 -O2 -DVERSION=\"2.5.3\" -DDOUBLE_PREC -I"$REF/utils" -I"$REF/theory/DD" -I"$REF/theory/DDrppi" -I"$REF/theory/DDsmu" \
      -I"$REF/theory/wp" -I"$REF/theory/xi" -I"$REF/mocks/DDtheta_mocks" caller.c -o _build/caller_ref $LINK
fi
gcc -std=c99 -O2 -DVERSION=\"2.5.3\" -DDOUBLE_PREC -I"$ROOT/include" caller.c -o _build/caller_inc $LINK
# The reference's own CPython extension modules, UNMODIFIED (theory/python_bindings/_countpairs.c:1136-2520,
# mocks/python_bindings/_countpairs_mocks.c), compiled where they lie and linked against libcorrfunc_b200.so instead of
# the reference's static libraries.  -DOMEGA_SAFE: utils/macros.h:101 uses a macro it never defines (it defines OMEGA).
if [ -d "$REF/theory/python_bindings" ]; then
  PYI=$(python -c "import sysconfig; print(sysconfig.get_paths()['include'])")
  NPI=$(python -c "import numpy; print(numpy.get_include())")
  EXT="-std=c99 -O2 -fPIC -shared -DVERSION=\"2.5.3\" -DDOUBLE_PREC -DNDEBUG -DOMEGA_SAFE=\"omega\" -I$PYI -I$NPI -I$REF/utils"
  gcc $EXT -I"$REF/theory/DD" -I"$REF/theory/DDrppi" -I"$REF/theory/DDsmu" -I"$REF/theory/wp" -I"$REF/theory/xi" -I"$REF/theory/vpf" \
      "$REF/theory/python_bindings/_countpairs.c" -o _build/_countpairs.so $LINK
  gcc $EXT -I"$REF/mocks/DDrppi_mocks" -I"$REF/mocks/DDsmu_mocks" -I"$REF/mocks/DDtheta_mocks" -I"$REF/mocks/vpf_mocks" \
      "$REF/mocks/python_bindings/_countpairs_mocks.c" -o _build/_countpairs_mocks.so $LINK
fi
