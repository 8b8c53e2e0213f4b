#!/bin/bash
# Builds tuning variants of the library into corrfunc_b200/csrc/variants/ (selected with CORRFUNC_B200_LIBPATH).
set -e
cd "$(dirname "$0")/../corrfunc_b200/csrc"
mkdir -p variants
build() { # name flags
  rm -f cuda/pairs_fast.o
  make -j8 VARIANT_FLAGS="$2" > /dev/null
  cp libcorrfunc_b200.so variants/libcorrfunc_b200_$1.so
}
build spi2 "-DFAST_SPI=2"
build pac "-DFAST_PA_COARSE"
build spi2pac "-DFAST_SPI=2 -DFAST_PA_COARSE"
rm -f cuda/pairs_fast.o
make -j8 > /dev/null
