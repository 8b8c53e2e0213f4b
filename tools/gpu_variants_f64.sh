#!/bin/bash
# double-precision fast kernel variants on c1, c2, c4 and config 5 in float64 (25 M points, same density)
for v in default "$@"; do
  if [ $v = default ]; then unset CORRFUNC_B200_LIBPATH; else export CORRFUNC_B200_LIBPATH=$PWD/corrfunc_b200/csrc/variants/libcorrfunc_b200_$v.so; fi
  for c in c1 c2 c4; do
    timeout 300 python bench.py --config $c --steps 3 --no-cpu-baseline 2>&1 | tail -1 | python tools/bench_summary.py "$v $c" | cut -c1-120
  done
  timeout 300 python bench.py --config c5d --npart 12000000 --same-density --steps 1 --no-cpu-baseline 2>&1 | tail -1 | python tools/bench_summary.py "$v c5d12M" | cut -c1-120
done
