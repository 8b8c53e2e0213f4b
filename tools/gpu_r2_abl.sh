#!/bin/bash
# ablations / block shapes of the sum kernel on the c3-like 3 M run (timing only)
for v in default "$@"; do
  if [ $v = default ]; then unset CORRFUNC_B200_LIBPATH; else export CORRFUNC_B200_LIBPATH=$PWD/corrfunc_b200/csrc/variants/libcorrfunc_b200_$v.so; fi
  echo "== $v"
  python tools/exp_sum.py 3e6 2>&1 | grep -v legacy | tail -7 | head -4
done
