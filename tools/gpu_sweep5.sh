#!/bin/bash
timeout 300 python -m pytest tests -m gpu -x -q -k "fast or brute or duplicate or narrow" 2>&1 | tail -2
for mb in 4 3; do
  CORRFUNC_B200_MINB=$mb timeout 300 python bench.py --config c5 --npart 6000000 --same-density --steps 2 --no-cpu-baseline 2>&1 | tail -1 | python tools/bench_summary.py "minb=$mb"
done
timeout 300 python bench.py --config c1 --steps 3 --no-cpu-baseline 2>&1 | tail -1 | python tools/bench_summary.py "c1"
timeout 300 python bench.py --config c4 --steps 1 --no-cpu-baseline 2>&1 | tail -1 | python tools/bench_summary.py "c4"
timeout 900 python bench.py --config c5 --steps 1 2>&1 | tail -1 | tee gpurun_out/bench_c5_full_v2.json | python tools/bench_summary.py "c5 full"
