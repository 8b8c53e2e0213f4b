#!/bin/bash
# round-1 GPU pass C: persistent-warp kernel, register-budget variants
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "fast or brute or duplicate or narrow or golden" 2>&1 | tail -2
for v in default p4l4 p6l4 p4l8 p5l8; do
  if [ $v = default ]; then unset CORRFUNC_B200_LIBPATH; else export CORRFUNC_B200_LIBPATH=$PWD/corrfunc_b200/csrc/variants/libcorrfunc_b200_$v.so; fi
  timeout 300 python bench.py --config c5 --npart 10000000 --same-density --steps 2 --no-cpu-baseline 2>&1 | tail -1 | python tools/bench_summary.py "$v"
done
unset CORRFUNC_B200_LIBPATH
timeout 300 python bench.py --config c1 --steps 3 --no-cpu-baseline 2>&1 | tail -1 | python tools/bench_summary.py "c1"
timeout 300 python bench.py --config c4 --steps 2 --no-cpu-baseline 2>&1 | tail -1 | python tools/bench_summary.py "c4"
