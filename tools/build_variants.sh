#!/bin/bash
# Builds tuning variants of the library into corrfunc_b200/csrc/variants/ (selected with CORRFUNC_B200_LIBPATH).
# usage: tools/build_variants.sh name1 "flags1" name2 "flags2" ...
set -e
cd "$(dirname "$0")/../corrfunc_b200/csrc"
mkdir -p variants
while [ $# -ge 2 ]; do
  rm -f cuda/pairs_fast.o cuda/pairs_generic.o cuda/pairs_sum.o
  make -j8 VARIANT_FLAGS="$2" > /dev/null
  cp libcorrfunc_b200.so variants/libcorrfunc_b200_$1.so
  grep -A3 "k_pairs_fastIfLi0ELb0ELb0" cuda/pairs_fast.o.ptxas.log | grep -E "Used|spill" | tr '\n' ' '; echo " <- $1"
  shift 2
done
rm -f cuda/pairs_fast.o cuda/pairs_generic.o cuda/pairs_sum.o
make -j8 > /dev/null
