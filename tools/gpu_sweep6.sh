#!/bin/bash
for v in default spi2 pac spi2pac; do
  if [ $v = default ]; then unset CORRFUNC_B200_LIBPATH; else export CORRFUNC_B200_LIBPATH=$PWD/corrfunc_b200/csrc/variants/libcorrfunc_b200_$v.so; fi
  for mb in 4 3; do
    CORRFUNC_B200_MINB=$mb timeout 300 python bench.py --config c5 --npart 6000000 --same-density --steps 2 --no-cpu-baseline 2>&1 | tail -1 | python tools/bench_summary.py "$v minb=$mb"
  done
done
