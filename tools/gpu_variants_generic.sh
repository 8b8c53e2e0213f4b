#!/bin/bash
# generic-kernel variants (named on the command line) on configs 3 and 2-DDrppi
for v in default "$@"; do
  if [ $v = default ]; then unset CORRFUNC_B200_LIBPATH; else export CORRFUNC_B200_LIBPATH=$PWD/corrfunc_b200/csrc/variants/libcorrfunc_b200_$v.so; fi
  for c in c3 c2rppi c2rppi32; do
    timeout 300 python bench.py --config $c --steps 3 --no-cpu-baseline 2>&1 | tail -1 | python tools/bench_summary.py "$v $c" | cut -c1-110
  done
done
