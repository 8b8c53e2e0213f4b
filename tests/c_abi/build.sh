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
