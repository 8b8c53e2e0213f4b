#!/bin/bash
for mb in 4 5 6; do
  CORRFUNC_B200_MINB=$mb timeout 300 python bench.py --config c5 --npart 6000000 --same-density --steps 2 --no-cpu-baseline 2>&1 | tail -1 | python tools/bench_summary.py "minb=$mb"
done
for mb in 3 4 5; do
  CORRFUNC_B200_MINB=$mb timeout 300 python bench.py --config c4 --steps 1 --no-cpu-baseline 2>&1 | tail -1 | python tools/bench_summary.py "c4 minb=$mb"
  CORRFUNC_B200_MINB=$mb timeout 300 python bench.py --config c1 --steps 3 --no-cpu-baseline 2>&1 | tail -1 | python tools/bench_summary.py "c1 minb=$mb"
done
