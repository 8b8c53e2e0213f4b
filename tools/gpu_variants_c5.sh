#!/bin/bash
# full-size config 5 (and c1/c4 spot checks) for the library variants named on the command line
timeout 600 python -m pytest tests -m gpu -x -q -k "fast or brute or duplicate or narrow or golden" 2>&1 | tail -1
for v in default "$@"; do
  if [ $v = default ]; then unset CORRFUNC_B200_LIBPATH; else export CORRFUNC_B200_LIBPATH=$PWD/corrfunc_b200/csrc/variants/libcorrfunc_b200_$v.so; fi
  timeout 300 python bench.py --config c5 --npart 25000000 --same-density --steps 2 --no-cpu-baseline 2>&1 | tail -1 | python tools/bench_summary.py "$v c5sd25M"
  timeout 300 python bench.py --config c4 --steps 2 --no-cpu-baseline 2>&1 | tail -1 | python tools/bench_summary.py "$v c4"
done
