#!/bin/bash
# resident-block cap x loop-body size (instruction-cache hypothesis), config 5 at 25 M points / same density
for v in default spioff; do
  if [ $v = default ]; then unset CORRFUNC_B200_LIBPATH; else export CORRFUNC_B200_LIBPATH=$PWD/corrfunc_b200/csrc/variants/libcorrfunc_b200_$v.so; fi
  for b in 2 3 4; do
    CORRFUNC_B200_FAST_BLOCKS_PER_SM=$b timeout 300 python bench.py --config c5 --npart 25000000 --same-density --steps 2 --no-cpu-baseline 2>&1 | tail -1 | python tools/bench_summary.py "$v blocks=$b"
  done
done
