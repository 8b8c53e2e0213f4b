#!/bin/bash
# quick regression + speed check of the fast kernel (default library, or CORRFUNC_B200_LIBPATH variants given as args)
timeout 600 python -m pytest tests -m gpu -x -q -k "fast or brute or duplicate or narrow or golden or full_size" 2>&1 | tail -2
for v in default "$@"; do
  if [ $v = default ]; then unset CORRFUNC_B200_LIBPATH; else export CORRFUNC_B200_LIBPATH=$PWD/corrfunc_b200/csrc/variants/libcorrfunc_b200_$v.so; fi
  timeout 300 python bench.py --config c5 --npart 10000000 --same-density --steps 2 --no-cpu-baseline 2>&1 | tail -1 | python tools/bench_summary.py "$v c5sd"
  timeout 300 python bench.py --config c1 --steps 3 --no-cpu-baseline 2>&1 | tail -1 | python tools/bench_summary.py "$v c1"
  timeout 300 python bench.py --config c4 --steps 2 --no-cpu-baseline 2>&1 | tail -1 | python tools/bench_summary.py "$v c4"
  timeout 300 python bench.py --config c2 --steps 3 --no-cpu-baseline 2>&1 | tail -1 | python tools/bench_summary.py "$v c2"
done
