#!/bin/bash
# bench lines of the sum-kernel configs for the default library and the variants given as arguments (no parity suite)
for v in default "$@"; do
  if [ $v = default ]; then unset CORRFUNC_B200_LIBPATH; else export CORRFUNC_B200_LIBPATH=$PWD/corrfunc_b200/csrc/variants/libcorrfunc_b200_$v.so; fi
  echo "== $v"
  for c in c3 m1 m2; do
    timeout 600 python bench.py --config $c --steps 3 --no-cpu-baseline 2>/dev/null | python tools/bench_summary.py $c | cut -c1-200
  done
  python tools/exp_sum.py 3e6 2>&1 | grep -v legacy | tail -7
done
